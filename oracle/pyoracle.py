"""ctypes door onto the parity checkers -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

* ``Oracle``  : oracle/_build/libkworacle.so, our plain-C restatement (oracle/fd1d_oracle.c).
* ``RefLib``  : oracle/_ref/libkwref.so, the UNMODIFIED reference compiled from
                /root/reference/src (oracle/Makefile, target ``ref``); may be absent.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product package (kwinto-cuda_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libkworacle.so")
REF_SO = os.path.join(HERE, "_ref", "libkwref.so")

# layout of kw::Option, /root/reference/src/Core/kwAsset.h:12-22 (56 B)
OPTION_DTYPE = np.dtype(
    {
        "names": ["t", "k", "z", "r", "q", "s", "e", "w"],
        "formats": ["<f8", "<f8", "<f8", "<f8", "<f8", "<f8", "u1", "i1"],
        "offsets": [0, 8, 16, 24, 32, 40, 48, 49],
        "itemsize": 56,
    }
)


def build(ref: bool = True) -> None:
    """Compile the checker libraries (gcc only).  ``ref`` is attempted only where the
    reference tree is mounted; on the GPU box the prebuilt oracle/_ref travels with the repo."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def _as_options(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=OPTION_DTYPE)
    assert a.dtype.itemsize == 56
    return a


class Oracle:
    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        self.lib = C.CDLL(path)
        L = self.lib
        L.kwo_fd1d_price.restype = C.c_int
        L.kwo_fd1d_price.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_int64, C.c_int64,
                                     C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint32), C.c_char_p, C.c_size_t]
        L.kwo_fd1d_bs_price.restype = C.c_int
        L.kwo_fd1d_bs_price.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_int64, C.c_int64,
                                        C.c_int, C.c_int, C.c_void_p, C.c_char_p, C.c_size_t]
        L.kwo_bs_price.restype = None
        L.kwo_bs_price.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.kwo_fd1d_solution.restype = C.c_int
        L.kwo_fd1d_solution.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.kwo_solve_tridiagonal.restype = C.c_int
        L.kwo_solve_tridiagonal.argtypes = [C.c_int] + [C.c_void_p] * 6
        L.kwo_x_grid.restype = None
        L.kwo_x_grid.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_void_p]
        L.kwo_max_threads.restype = C.c_int
        L.kwo_sizeof_option.restype = C.c_size_t
        assert L.kwo_sizeof_option() == 56

    def max_threads(self) -> int:
        return int(self.lib.kwo_max_threads())

    def fd1d(self, options, tdim=512, xdim=512, density=0.25, scale=50.0, compress=True, nthreads=0,
             mode="FD1D"):
        """-> (prices, error string).  Mirrors Fd1d_Pricer::price (src/Pricer/kwFd1d.cpp:21)."""
        o = _as_options(options)
        n = o.shape[0]
        prices = np.full(n, np.nan)
        err = C.create_string_buffer(256)
        if mode == "FD1D":
            npde = C.c_uint32(0)
            rc = self.lib.kwo_fd1d_price(o.ctypes.data, n, density, scale, tdim, xdim, int(compress), nthreads,
                                         prices.ctypes.data, C.byref(npde), err, 256)
        elif mode == "FD1D-BS":
            rc = self.lib.kwo_fd1d_bs_price(o.ctypes.data, n, density, scale, tdim, xdim, int(compress),
                                            nthreads, prices.ctypes.data, err, 256)
        elif mode == "BS":
            self.lib.kwo_bs_price(o.ctypes.data, n, prices.ctypes.data)
            rc = 0
        else:
            raise ValueError(mode)
        return prices, (err.value.decode() if rc else "")

    def set_libm_jitter(self, ulps: int) -> None:
        """Test hook of fd1d_oracle.c: move every sinh() / exp() result by 0..ulps ulp (0 = the reference's bits)."""
        self.lib.kwo_set_libm_jitter.restype = None
        self.lib.kwo_set_libm_jitter.argtypes = [C.c_int]
        self.lib.kwo_set_libm_jitter(int(ulps))

    def libm_sensitivity(self, options, tdim, xdim, ulps=2, **kw):
        """max |price(jitter = ulps) - price(jitter = 0)| per option: how far two faithful builds of the reference
        against different libms may be apart on this grid shape (see fd1d_oracle.c: libm_jitter)."""
        base, err0 = self.fd1d(options, tdim, xdim, **kw)
        try:
            self.set_libm_jitter(ulps)
            jit, err1 = self.fd1d(options, tdim, xdim, **kw)
        finally:
            self.set_libm_jitter(0)
        assert err0 == err1 == ""
        return np.abs(jit - base)

    def solution(self, option, tdim, xdim, density=0.25, scale=50.0):
        o = _as_options(np.asarray(option).reshape(1))
        x = np.empty(xdim)
        v = np.empty(xdim)
        rc = self.lib.kwo_fd1d_solution(o.ctypes.data, density, scale, tdim, xdim, x.ctypes.data, v.ctypes.data)
        assert rc == 0
        return x, v

    def x_grid(self, z, t, xdim, density=0.25, scale=50.0):
        x = np.empty(xdim)
        self.lib.kwo_x_grid(z, t, density, scale, xdim, x.ctypes.data)
        return x

    def solve_tridiagonal(self, al, a, au, y):
        al, a, au, y = (np.ascontiguousarray(v, dtype=np.float64) for v in (al, a, au, y))
        n = a.shape[0]
        x = np.zeros(n)
        gam = np.zeros(n)
        rc = self.lib.kwo_solve_tridiagonal(n, al.ctypes.data, a.ctypes.data, au.ctypes.data, y.ctypes.data,
                                            x.ctypes.data, gam.ctypes.data)
        return x, rc


class RefLib:
    """The unmodified reference (its own thread pool, all host cores)."""

    def __init__(self, path: str = REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        L = self.lib
        L.kwref_price.restype = C.c_int
        L.kwref_price.argtypes = [C.c_char_p, C.c_double, C.c_double, C.c_longlong, C.c_longlong, C.c_void_p,
                                  C.c_size_t, C.c_void_p, C.c_char_p, C.c_size_t]
        L.kwref_pool_size.restype = C.c_int
        L.kwref_sizeof_option.restype = C.c_size_t
        assert L.kwref_sizeof_option() == 56

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def pool_size(self) -> int:
        return int(self.lib.kwref_pool_size())

    def price(self, options, tdim=512, xdim=512, density=0.25, scale=50.0, mode="FD1D"):
        o = _as_options(options)
        n = o.shape[0]
        prices = np.full(n, np.nan)
        err = C.create_string_buffer(512)
        rc = self.lib.kwref_price(mode.encode(), density, scale, tdim, xdim, o.ctypes.data, n,
                                  prices.ctypes.data, err, 512)
        return prices, (err.value.decode() if rc else "")
