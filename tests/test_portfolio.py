"""The callers either side of the hot path (SURVEY.md 8(f)-2,3): kw::Portfolio's CSV(+zstd) loader and
price statistics, in both host mirrors (Python kwfd1d.portfolio, C++ host/kw/kwPortfolio.h behind the
kwinto-gpu CLI).  Inputs are rebuilt from the committed golden vectors (tests/golden/*.npz, made from the
reference's own fixtures by tests/golden/make_golden.py) -- /root/reference is never read here."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden

CLI = os.path.join(ROOT, "kwinto-cuda_b200", "bin", "kwinto-gpu")
# the reference's own output for `kwinto price -p FD1D test/portfolio_fd1d.csv` at 512^2
REF_STATS_FD1D = {"rmse": 9.883494e-04, "rrmse": 8.115662e-05, "mae": 9.319823e-03, "mre": 1.031885e-03, "total": 4543}


def write_csv(path, options, prices, zst=False):
    """The fixture format (test/portfolio.py:97): shortest round-trip floats, so parsing is exact."""
    lines = ["expiry,spot,strike,volatility,interest_rate,dividend_rate,parity,exercise,price"]
    for o, p in zip(options, prices):
        lines.append(",".join([repr(float(o["t"])), repr(float(o["s"])), repr(float(o["k"])), repr(float(o["z"])),
                               repr(float(o["r"])), repr(float(o["q"])), "c" if o["w"] > 0 else "p",
                               "a" if o["e"] else "e", repr(float(p))]))
    text = ("\n".join(lines) + "\n").encode()
    if zst:
        import pyarrow as pa

        with pa.output_stream(path, compression="zstd") as f:
            f.write(text)
    else:
        with open(path, "wb") as f:
            f.write(text)


@pytest.fixture(scope="module")
def csv_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("portfolio")
    g = load_golden("portfolio_fd1d")
    plain, zst = str(d / "portfolio_fd1d.csv"), str(d / "portfolio_fd1d.csv.zst")
    write_csv(plain, g["options"], g["quantlib"])
    write_csv(zst, g["options"], g["quantlib"], zst=True)
    return plain, zst, g


def test_python_loader_round_trip(csv_files):
    from kwfd1d.portfolio import Portfolio

    plain, zst, g = csv_files
    for path in (plain, zst):
        p = Portfolio()
        assert p.load(path) == ""
        assert p.assets.shape == (6000,)
        for f in ("t", "k", "z", "r", "q", "s", "e", "w"):
            assert np.array_equal(p.assets[f], g["options"][f]), f
        assert np.array_equal(p.prices, g["quantlib"])
    assert Portfolio().load("/no/such/file.csv").startswith("Portfolio::load : Failed to open")


def test_loader_rejects_missing_columns(tmp_path):
    from kwfd1d.portfolio import Portfolio

    bad = tmp_path / "bad.csv"
    bad.write_text("expiry,spot,strike\n1,2,3\n")
    err = Portfolio().load(str(bad))
    assert err.startswith("Portfolio::load : Some option data is missing: e=-1, k=2")


def test_price_statistics_reproduce_the_reference_run(csv_files):
    """printPricesStats over the reference's own 512^2 prices reproduces the block the reference prints."""
    from kwfd1d.portfolio import Portfolio, format_stats

    plain, _, g = csv_files
    p = Portfolio()
    assert p.load(plain) == ""
    st = p.stats(g["fd1d_512"])
    assert st["total"] == REF_STATS_FD1D["total"]
    for k in ("rmse", "rrmse", "mae", "mre"):
        assert abs(st[k] - REF_STATS_FD1D[k]) <= 1e-6 * REF_STATS_FD1D[k], (k, st[k])
    txt = format_stats(st)
    assert "RMSE : 9.883494e-04" in txt and "total : 4543 options" in txt and "MAE Asset : <Option s=" in txt


def parse_stats(out):
    m = {k: float(re.search(k.upper() + r" : (\S+)", out).group(1)) for k in ("rmse", "rrmse", "mae", "mre")}
    m["total"] = int(re.search(r"total : (\d+) options", out).group(1))
    return m


def test_cli_loads_plain_and_zst(csv_files):
    """No GPU needed: the C++ loader runs first and prints the asset count; without a device the pricer
    then fails loudly (there is no CPU fallback)."""
    assert os.path.exists(CLI), "run __graft_entry__.build() first"
    plain, zst, _ = csv_files
    for path in (plain, zst):
        r = subprocess.run([CLI, "price", "-p", "FD1D", "-t", "64", "-x", "64", path], capture_output=True, text=True,
                           timeout=300)
        assert "Assets : 6000" in r.stdout, r.stdout + r.stderr
        if r.returncode != 0:
            assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr, r.stderr
    r = subprocess.run([CLI, "price", "/no/such.csv"], capture_output=True, text=True)
    assert r.returncode == 1 and "Portfolio::load : Failed to open" in r.stderr
    r = subprocess.run([CLI, "frobnicate", "x.csv"], capture_output=True, text=True)
    assert r.returncode == 1 and "expected a command" in r.stderr


@pytest.mark.gpu
def test_cli_price_matches_reference_statistics(csv_files):
    """`kwinto-gpu price -p FD1D <fixture>` prints the reference's statistics block (config 1, 512^2)."""
    _, zst, _ = csv_files
    r = subprocess.run([CLI, "price", "-p", "FD1D", "-v", zst], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    st = parse_stats(r.stdout)
    assert st["total"] == REF_STATS_FD1D["total"]
    for k in ("rmse", "rrmse", "mae", "mre"):
        assert abs(st[k] - REF_STATS_FD1D[k]) <= 2e-6 * REF_STATS_FD1D[k], (k, st[k], r.stdout)
    # the CLI default is the control-variate pricer (src/kwinto.cpp:28)
    r = subprocess.run([CLI, "price", zst], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "Price Statistics" in r.stdout


@pytest.mark.gpu
def test_cli_bench_blocks(csv_files):
    """The historic benchmark output (log/z800_1024_32768.log): devices, portfolio, timing and error blocks."""
    plain, _, _ = csv_files
    r = subprocess.run([CLI, "bench", "-v", "--gpu32", "--gpu64", "--put", "-b", "4096", "-n", "3", "-x", "256", "-t",
                        "256", plain], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    for needle in ("Devices Info #0", "Total SM:", "Peak Bandwidth:", "Batch count: 3", "Batch size : 4096",
                   "Benchmark for Fd1dGpu_Pricer<float>::price", "Benchmark for Fd1dGpu_Pricer<double>::price",
                   "funCall : 3 times", "avgTime :", "Errors for Fd1dGpu_Pricer<double>::price", "RRMSE :"):
        assert needle in out, (needle, out)


@pytest.mark.gpu
def test_python_portfolio_price(csv_files):
    import kwfd1d
    from kwfd1d.portfolio import Portfolio

    _, zst, g = csv_files
    p = Portfolio()
    assert p.load(zst) == ""
    cfg = kwfd1d.Config(PRICER="FD1D-GPU")
    err, prices = p.price(cfg)
    assert err == "" and float(np.max(np.abs(prices - g["fd1d_512"]))) <= 1e-9
    st = p.stats(prices)
    assert st["total"] == 4543 and abs(st["rmse"] - REF_STATS_FD1D["rmse"]) <= 1e-6 * REF_STATS_FD1D["rmse"]
    err, _ = p.price(kwfd1d.Config(PRICER="NOPE"))
    assert err == "Portfolio::price : PricerFactory: Unknown PRICER = NOPE"


def test_cli_bench_cpu_arms(csv_files):
    """The --cpu64 / --cpu32 arms of the historic bench (log/z800_1024_32768.log:42-78, Makefile:5).  The CPU arm is the
    reference's own Fd1d_Pricer, compiled from the reference tree into the DRIVER binary kwinto-gpu-ref (never into
    libkwfd1d.so), so it runs without a GPU; the plain driver refuses the flag."""
    plain, _, _ = csv_files
    r = subprocess.run([CLI, "bench", "--cpu64", "-b", "64", "-n", "2", "-x", "64", "-t", "64", plain],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "kwinto-gpu-ref" in r.stderr
    ref_cli = CLI + "-ref"
    if not os.path.exists(ref_cli):
        pytest.skip("kwinto-gpu-ref is built only where the reference tree is mounted")
    r = subprocess.run([ref_cli, "bench", "--cpu32", "--cpu64", "-p", "FD1D", "--put", "-b", "256", "-n", "2", "-x", "128",
                        "-t", "128", plain], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    for needle in ("Batch count: 2", "Batch size : 256", "Benchmark for Fd1d_Pricer<float>::price", "not available",
                   "Benchmark for Fd1d_Pricer<double>::price", "funCall : 2 times", "options/s :",
                   "Errors for Fd1d_Pricer<double>::price", "RRMSE :"):
        assert needle in out, (needle, out)
    rrmse = float(re.search(r"RRMSE : (\S+)", out).group(1))
    assert 0 < rrmse < 5e-2  # 128 x 128 grid against the QuantLib reference prices
