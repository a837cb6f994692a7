"""Synthetic option batches for the benchmark configs (SURVEY.md 8(d) "Synthetic generator").

Every option gets its own (t, z, r, q) so that one option = one PDE and the pricer's chain
compression (reference: src/Pricer/kwFd1d.cpp:28-65) cannot shortcut the work.

The stream is std::mt19937_64 (so the C++ driver in host/ draws the same portfolio with
``std::mt19937_64 rng(seed); u = (rng() >> 11) * 0x1p-53``), implemented here in numpy with a
vectorised twist.  Draw order per option: t, z, r, q, k.
"""
from __future__ import annotations

import numpy as np

from .types import OPTION_DTYPE

_NN, _MM = 312, 156
_MATRIX_A = np.uint64(0xB5026F5AA96619E9)
_UM = np.uint64(0xFFFFFFFF80000000)
_LM = np.uint64(0x7FFFFFFF)
_ONE = np.uint64(1)


class MT19937_64:
    """Bit-exact std::mt19937_64 (10000th output of seed 5489 is 9981545732273789042)."""

    def __init__(self, seed: int = 5489):
        mt = np.empty(_NN, dtype=np.uint64)
        x = seed & 0xFFFFFFFFFFFFFFFF
        mt[0] = x
        for i in range(1, _NN):
            x = (6364136223846793005 * (x ^ (x >> 62)) + i) & 0xFFFFFFFFFFFFFFFF
            mt[i] = x
        self.mt = mt
        self.buf = np.empty(0, dtype=np.uint64)

    def _twist(self) -> None:
        mt = self.mt

        def mix(cur, nxt, far):
            x = (cur & _UM) | (nxt & _LM)
            return far ^ (x >> _ONE) ^ ((x & _ONE) * _MATRIX_A)

        # i in [0, 156): needs old mt[i], old mt[i+1], old mt[i+156]
        mt[0:_MM] = mix(mt[0:_MM], mt[1:_MM + 1], mt[_MM:_NN])
        # i in [156, 311): needs old mt[i], old mt[i+1], NEW mt[i-156]
        mt[_MM:_NN - 1] = mix(mt[_MM:_NN - 1], mt[_MM + 1:_NN], mt[0:_MM - 1])
        # i = 311: old mt[311], NEW mt[0], NEW mt[155]
        mt[_NN - 1:_NN] = mix(mt[_NN - 1:_NN], mt[0:1], mt[_MM - 1:_MM])

    @staticmethod
    def _temper(x):
        x = x ^ ((x >> np.uint64(29)) & np.uint64(0x5555555555555555))
        x = x ^ ((x << np.uint64(17)) & np.uint64(0x71D67FFFEDA60000))
        x = x ^ ((x << np.uint64(37)) & np.uint64(0xFFF7EEE000000000))
        x = x ^ (x >> np.uint64(43))
        return x

    def raw(self, n: int) -> np.ndarray:
        out = [self.buf]
        have = self.buf.shape[0]
        while have < n:
            self._twist()
            out.append(self._temper(self.mt.copy()))
            have += _NN
        allv = np.concatenate(out)
        self.buf = allv[n:]
        return allv[:n]

    def uniform(self, n: int) -> np.ndarray:
        """(rng() >> 11) * 2^-53 in [0, 1)."""
        return (self.raw(n) >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def synthetic_options(n: int, seed: int, european_every: int = 0, call_every: int = 0) -> np.ndarray:
    """n American puts with unique chains: t in [1/12, 2], z in [.1,.6], r in [.02,.1], q in [0,.12],
    k in [50,150], s = 100 (ranges of the reference's test/portfolio.py:66-91).  ``european_every`` /
    ``call_every`` > 0 turn every k-th option European / into a call (mix variant, not the headline)."""
    u = MT19937_64(seed).uniform(5 * n).reshape(n, 5)
    o = np.zeros(n, dtype=OPTION_DTYPE)
    o["t"] = 1.0 / 12 + u[:, 0] * (2.0 - 1.0 / 12)
    o["z"] = 0.1 + 0.5 * u[:, 1]
    o["r"] = 0.02 + 0.08 * u[:, 2]
    o["q"] = 0.12 * u[:, 3]
    o["k"] = 50.0 + 100.0 * u[:, 4]
    o["s"] = 100.0
    o["e"] = 1
    o["w"] = -1
    if european_every > 0:
        o["e"][european_every - 1::european_every] = 0
    if call_every > 0:
        o["w"][call_every - 1::call_every] = 1
    return o
