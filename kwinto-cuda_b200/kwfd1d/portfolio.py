"""kw::Portfolio (src/Utils/kwPortfolio.h:12-36, src/Utils/kwPortfolio.cpp:13-163) for the Python host
side: CSV(+zstd) loading with the reference's column names, pricing through the GPU factory, and the
"Price Statistics" block.  The C++ twin is kwinto-cuda_b200/host/kw/kwPortfolio.h."""
from __future__ import annotations

import csv
import io
import math
from typing import Optional, Tuple

import numpy as np

from .types import OPTION_DTYPE

COLUMNS = {"exercise": "e", "strike": "k", "dividend_rate": "q", "interest_rate": "r", "spot": "s",
           "expiry": "t", "price": "v", "parity": "w", "volatility": "z"}


def _read_text(path: str) -> str:
    if path.endswith(".zst"):
        import pyarrow as pa  # the image has no zstd module; pyarrow bundles the codec

        with pa.input_stream(path, compression="zstd") as f:
            return f.read().decode()
    with open(path, "r", newline="") as f:
        return f.read()


class Portfolio:
    def __init__(self):
        self.assets = np.zeros(0, dtype=OPTION_DTYPE)
        self.prices = np.zeros(0, dtype=np.float64)

    def load(self, path: str) -> str:
        """Portfolio::load (src/Utils/kwPortfolio.cpp:13-84); *.zst is decoded first."""
        try:
            text = _read_text(path)
        except OSError:
            return "Portfolio::load : Failed to open " + path
        rows = csv.reader(io.StringIO(text))
        header = next(rows, [])
        idx = {short: -1 for short in COLUMNS.values()}
        for i, name in enumerate(header):
            if name in COLUMNS:
                idx[COLUMNS[name]] = i
        if min(idx.values()) < 0:
            return ("Portfolio::load : Some option data is missing: " +
                    ", ".join(f"{k}={idx[k]}" for k in ("e", "k", "q", "r", "s", "t", "v", "w", "z")))
        recs, prices = [], []
        for vals in rows:
            if not vals:
                continue
            recs.append((float(vals[idx["t"]]), float(vals[idx["k"]]), float(vals[idx["z"]]), float(vals[idx["r"]]),
                         float(vals[idx["q"]]), float(vals[idx["s"]]), 1 if vals[idx["e"]] == "a" else 0,
                         1 if vals[idx["w"]] == "c" else -1))
            prices.append(float(vals[idx["v"]]))
        self.assets = np.array(recs, dtype=OPTION_DTYPE)
        self.prices = np.array(prices, dtype=np.float64)
        return ""

    def price(self, config) -> Tuple[str, Optional[np.ndarray]]:
        """Portfolio::price (src/Utils/kwPortfolio.cpp:87-98) on the GPU factory."""
        from . import PricerFactory

        err, pricer = PricerFactory.create(config)
        if err:
            return "Portfolio::price : " + err, None
        err, prices = pricer.price(self.assets)
        if err:
            return "Portfolio::price : " + err, prices
        return "", prices

    def stats(self, prices: np.ndarray, tolerance: float = 0.5) -> dict:
        """The arithmetic of Portfolio::printPricesStats (src/Utils/kwPortfolio.cpp:101-146)."""
        return price_stats(self.prices, prices, tolerance, self.assets)

    def print_prices_stats(self, prices: np.ndarray, tolerance: float = 0.5) -> str:
        return format_stats(self.stats(prices, tolerance))


def price_stats(want: np.ndarray, got: np.ndarray, tolerance: float = 0.5, assets=None) -> dict:
    keep = ~(want < tolerance)
    w, g = want[keep], got[keep]
    ad = np.abs(w - g)
    rd = ad / w
    n = int(keep.sum())
    am, rm = ad.sum() / n, rd.sum() / n
    out = {"rmse": math.sqrt((ad * ad).sum() / n - am * am), "rrmse": math.sqrt((rd * rd).sum() / n - rm * rm),
           "mae": float(ad.max()), "mre": float(rd.max()), "total": n}
    if assets is not None:
        a = assets[keep]
        out["mae_asset"], out["mre_asset"] = a[int(ad.argmax())], a[int(rd.argmax())]
    return out


def option_as_string(o) -> str:
    """operator<<(ostream, Option), src/Core/kwAsset.cpp:3-8 (default ostream float formatting: %g)."""
    return ("<Option s=%g t=%g, k=%g, z=%g, r=%g, q=%g, %s, %s>" %
            (o["s"], o["t"], o["k"], o["z"], o["r"], o["q"], "amer" if o["e"] else "euro",
             "call" if o["w"] > 0 else "put"))


def format_stats(st: dict) -> str:
    lines = ["Price Statistics", "       RMSE : %.6e" % st["rmse"], "      RRMSE : %.6e" % st["rrmse"],
             "        MAE : %.6e" % st["mae"], "        MRE : %.6e" % st["mre"]]
    if "mae_asset" in st:
        lines += ["  MAE Asset : " + option_as_string(st["mae_asset"]), "  MRE Asset : " + option_as_string(st["mre_asset"])]
    lines += ["      total : %d options" % st["total"], ""]
    return "\n".join(lines)
