"""kwfd1d -- Python host side of the B200-native Fd1d pricer.

A thin ctypes binding over the C ABI (include/kw_fd1d.h, lib/libkwfd1d.so) that mirrors the
reference's C++ interface for this one path, with the reference's names and error behaviour:

    kw::Config         src/Core/kwConfig.h:21-73         -> Config
    kw::Pricer         src/Pricer/kwPricer.h:12-22       -> Pricer (init / price, "" = success)
    kw::PricerFactory  src/Pricer/kwPricerFactory.h:15-41 -> PricerFactory.create, keys
                       "FD1D-GPU" (the drop-in for "FD1D") and "FD1D-BS-GPU" (for "FD1D-BS")

The C++ twin of this file is kwinto-cuda_b200/host/kw/ (what a maintainer of the reference
would compile in).  There is no CPU fallback: if the CUDA library is
missing or there is no device, loading / init fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

from .types import OPTION_DTYPE, make_options  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libkwfd1d.so")

KW_FD1D_OK, KW_FD1D_EINVAL, KW_FD1D_ECUDA, KW_FD1D_ERANGE, KW_FD1D_ENOMEM = range(5)
LAYOUT_AUTO, LAYOUT_REG, LAYOUT_SOA = 0, 1, 2
LAYOUTS = {"auto": LAYOUT_AUTO, "reg": LAYOUT_REG, "soa": LAYOUT_SOA}
PRECISIONS = {"f64": 0, "f32": 1}


class _CConfig(C.Structure):  # struct kw_fd1d_config
    _fields_ = [("density", C.c_double), ("scale", C.c_double), ("t_grid_size", C.c_int64),
                ("x_grid_size", C.c_int64), ("device", C.c_int32), ("precision", C.c_int32),
                ("layout", C.c_int32), ("compress", C.c_int32), ("variant", C.c_int32),
                ("exact", C.c_int32), ("bs_fused", C.c_int32), ("n_devices", C.c_int32)]


class _CInfo(C.Structure):  # struct kw_fd1d_info
    _fields_ = [("device", C.c_int32), ("sm_count", C.c_int32), ("layout", C.c_int32), ("variant", C.c_int32),
                ("threads_per_pde", C.c_int32), ("nodes_per_thread", C.c_int32), ("ctas_per_sm", C.c_int32),
                ("regs_per_thread", C.c_int32), ("smem_per_cta", C.c_int32), ("grid", C.c_int32),
                ("sm_clock_khz", C.c_int32), ("launches", C.c_int32), ("last_kernel_ms", C.c_double),
                ("last_n_pde", C.c_uint64), ("mode_count", C.c_uint32 * 6), ("device_name", C.c_char * 128),
                ("n_devices", C.c_int32), ("devices_used", C.c_int32), ("last_wall_ms", C.c_double),
                ("long_chains", C.c_uint32), ("reserved_", C.c_uint32)]


_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen lib/libkwfd1d.so.  Raises if it has not been built (__graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the Fd1d GPU pricer)")
    L = C.CDLL(path)
    H = C.c_void_p
    L.kw_fd1d_config_default.argtypes = [C.POINTER(_CConfig)]
    L.kw_fd1d_config_default.restype = None
    L.kw_fd1d_create.argtypes = [C.POINTER(_CConfig), C.POINTER(H)]
    L.kw_fd1d_create.restype = C.c_int
    L.kw_fd1d_create_multi.argtypes = [C.POINTER(_CConfig), C.POINTER(C.c_int32), C.c_int32, C.POINTER(H)]
    L.kw_fd1d_create_multi.restype = C.c_int
    L.kw_fd1d_has_variant.argtypes = [C.c_int32, C.c_int32]
    L.kw_fd1d_has_variant.restype = C.c_int
    L.kw_fd1d_destroy.argtypes = [H]
    L.kw_fd1d_destroy.restype = None
    L.kw_fd1d_price.argtypes = [H, C.c_void_p, C.c_size_t, C.c_void_p]
    L.kw_fd1d_price.restype = C.c_int
    L.kw_fd1d_price_bs.argtypes = [H, C.c_void_p, C.c_size_t, C.c_void_p]
    L.kw_fd1d_price_bs.restype = C.c_int
    L.kw_fd1d_price_device.argtypes = [H, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.kw_fd1d_price_device.restype = C.c_int
    L.kw_fd1d_sync.argtypes = [H, C.c_void_p]
    L.kw_fd1d_sync.restype = C.c_int
    L.kw_fd1d_last_error.argtypes = [H]
    L.kw_fd1d_last_error.restype = C.c_char_p
    L.kw_fd1d_get_info.argtypes = [H, C.POINTER(_CInfo)]
    L.kw_fd1d_get_info.restype = C.c_int
    L.kw_fd1d_fp64_peak.argtypes = [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.kw_fd1d_fp64_peak.restype = C.c_int
    L.kw_fd1d_microbench.argtypes = [C.c_int32, C.POINTER(C.c_double)]
    L.kw_fd1d_microbench.restype = C.c_int
    L.kw_fd1d_tmem_probe.argtypes = [C.c_int32, C.POINTER(C.c_double)]
    L.kw_fd1d_tmem_probe.restype = C.c_int
    L.kw_fd1d_dfma_probe.argtypes = [C.c_int32, C.POINTER(C.c_double)]
    L.kw_fd1d_dfma_probe.restype = C.c_int
    L.kw_fd1d_version.restype = C.c_char_p
    _lib = L
    return L


EXPORTED_SYMBOLS = [  # every entry point include/kw_fd1d.h declares
    "kw_fd1d_config_default", "kw_fd1d_create", "kw_fd1d_create_multi", "kw_fd1d_destroy", "kw_fd1d_price", "kw_fd1d_price_device",
    "kw_fd1d_sync", "kw_fd1d_price_bs", "kw_fd1d_device_count", "kw_fd1d_get_device_props", "kw_fd1d_last_error", "kw_fd1d_get_info", "kw_fd1d_fp64_peak",
    "kw_fd1d_microbench", "kw_fd1d_tmem_probe", "kw_fd1d_dfma_probe", "kw_fd1d_has_variant", "kw_fd1d_version",
]


class Config:
    """kw::Config (src/Core/kwConfig.h:21-73): three typed maps; the getter is chosen by the TYPE of
    the default argument, so e.g. a density stored as an int is not seen by get(key, 0.25)."""

    def __init__(self, **kv):
        self._doubles, self._integers, self._strings = {}, {}, {}
        for k, v in kv.items():
            self.set(k, v)

    def set(self, key: str, value) -> None:
        if isinstance(value, bool):
            self._integers[key] = int(value)
        elif isinstance(value, (int, np.integer)):
            self._integers[key] = int(value)
        elif isinstance(value, (float, np.floating)):
            self._doubles[key] = float(value)
        elif isinstance(value, str):
            self._strings[key] = value
        else:
            raise TypeError(f"Config.set: unsupported type {type(value)}")

    def get(self, key: str, default):
        if isinstance(default, bool) or isinstance(default, (int, np.integer)):
            return self._integers.get(key, default)
        if isinstance(default, (float, np.floating)):
            return self._doubles.get(key, default)
        if isinstance(default, str):
            return self._strings.get(key, default)
        raise TypeError(f"Config.get: unsupported type {type(default)}")

    def erase(self, key: str) -> None:
        for m in (self._doubles, self._integers, self._strings):
            m.pop(key, None)


class Pricer:
    """kw::Pricer (src/Pricer/kwPricer.h:12-22).  Error = str, "" means success."""

    def init(self, config: Config) -> str:
        raise NotImplementedError

    def price(self, assets) -> Tuple[str, Optional[np.ndarray]]:
        """-> (error, prices).  The reference fills a caller-owned vector; here the array is returned."""
        raise NotImplementedError


class Fd1dGpu_Pricer(Pricer):
    """Drop-in for kw::Fd1d_Pricer (src/Pricer/kwFd1d.h:14-38) running on one B200.

    Keys (init): FD1D.DENSITY, FD1D.SCALE, FD1D.T_GRID_SIZE, FD1D.X_GRID_SIZE exactly as the
    reference (src/Pricer/kwFd1d.cpp:12-16) plus FD1D.GPU.DEVICE (int), FD1D.GPU.LAYOUT
    ("auto"|"reg"|"soa"), FD1D.GPU.PRECISION ("f64"), FD1D.GPU.COMPRESS (int 0/1),
    FD1D.GPU.DEVICES (int N or "0,1,..": one price() call over several GPUs), FD1D.GPU.VARIANT (int), FD1D.GPU.EXACT (int 0/1/2: 0 lets provably negligible carry terms be
    dropped, 2 keeps every term), FD1D.GPU.BS_FUSED (int, "FD1D-BS-GPU" only: 0 = fused
    American + European march (variant 257) for batches of a device wave or more, 1 = always two solves as
    the reference does, 4 = variant 257 for every batch size, 3 / 2 = the measured experiments 252 / 251)."""

    _mode_bs = False

    def __init__(self):
        self._h = C.c_void_p()
        self._lib = None

    def init(self, config: Config) -> str:
        self._lib = load_library()
        self.close()
        c = _CConfig()
        self._lib.kw_fd1d_config_default(C.byref(c))
        c.density = config.get("FD1D.DENSITY", 0.25)
        c.scale = config.get("FD1D.SCALE", 50.0)
        c.t_grid_size = config.get("FD1D.T_GRID_SIZE", 512)
        c.x_grid_size = config.get("FD1D.X_GRID_SIZE", 512)
        c.device = config.get("FD1D.GPU.DEVICE", 0)
        layout = config.get("FD1D.GPU.LAYOUT", "auto")
        prec = config.get("FD1D.GPU.PRECISION", "f64")
        if layout not in LAYOUTS:
            return f"Fd1dGpu_Pricer::init: unknown FD1D.GPU.LAYOUT = {layout}"
        if prec not in PRECISIONS:
            return f"Fd1dGpu_Pricer::init: unknown FD1D.GPU.PRECISION = {prec}"
        c.layout = LAYOUTS[layout]
        c.precision = PRECISIONS[prec]
        c.compress = config.get("FD1D.GPU.COMPRESS", 1)
        c.variant = config.get("FD1D.GPU.VARIANT", 0)
        c.exact = config.get("FD1D.GPU.EXACT", 0)
        c.bs_fused = config.get("FD1D.GPU.BS_FUSED", 0)
        # FD1D.GPU.DEVICES: an int N (devices DEVICE ... DEVICE + N - 1; -1 = all visible) or a comma-separated list
        # of ordinals ("0,1,2,3"; an ordinal may repeat): ONE price() call then uses all of them (kw_fd1d_create_multi)
        dev_list = config.get("FD1D.GPU.DEVICES", "")
        c.n_devices = config.get("FD1D.GPU.DEVICES", 0)
        h = C.c_void_p()
        if dev_list:
            try:
                devs = [int(d) for d in str(dev_list).split(",")]
            except ValueError:
                return f"Fd1dGpu_Pricer::init: FD1D.GPU.DEVICES = {dev_list} is not a list of device ordinals"
            arr = (C.c_int32 * len(devs))(*devs)
            rc = self._lib.kw_fd1d_create_multi(C.byref(c), arr, len(devs), C.byref(h))
        else:
            rc = self._lib.kw_fd1d_create(C.byref(c), C.byref(h))
        if rc != KW_FD1D_OK:
            msg = self._lib.kw_fd1d_last_error(h).decode() if h else "Fd1dGpu_Pricer::init: failed"
            if h:
                self._lib.kw_fd1d_destroy(h)
            return msg
        self._h = h
        return ""

    def _error(self) -> str:
        return self._lib.kw_fd1d_last_error(self._h).decode()

    def price(self, assets) -> Tuple[str, Optional[np.ndarray]]:
        if not self._h:
            return "Fd1dGpu_Pricer::price: pricer was not initialised", None
        a = np.ascontiguousarray(assets, dtype=OPTION_DTYPE)
        n = a.shape[0]
        if n == 0:
            return "", None  # src/Pricer/kwFd1d.cpp:24-26: prices untouched
        prices = np.empty(n, dtype=np.float64)
        fn = self._lib.kw_fd1d_price_bs if self._mode_bs else self._lib.kw_fd1d_price
        rc = fn(self._h, a.ctypes.data, n, prices.ctypes.data)
        if rc != KW_FD1D_OK:
            return self._error(), (prices if rc == KW_FD1D_ERANGE else None)
        return "", prices

    # -- device-resident path (bench: inputs already in HBM) ---------------------------------
    def price_device(self, d_assets_ptr: int, n: int, d_prices_ptr: int, stream_ptr: int = 0) -> str:
        rc = self._lib.kw_fd1d_price_device(self._h, C.c_void_p(d_assets_ptr), n, C.c_void_p(d_prices_ptr),
                                            C.c_void_p(stream_ptr))
        return "" if rc == KW_FD1D_OK else self._error()

    def sync(self, stream_ptr: int = 0) -> str:
        rc = self._lib.kw_fd1d_sync(self._h, C.c_void_p(stream_ptr))
        return "" if rc == KW_FD1D_OK else self._error()

    def info(self) -> dict:
        i = _CInfo()
        self._lib.kw_fd1d_get_info(self._h, C.byref(i))
        d = {f[0]: getattr(i, f[0]) for f in _CInfo._fields_ if f[0] != "reserved"}
        d["device_name"] = i.device_name.decode()
        d["mode_count"] = list(i.mode_count)[:6]
        d["layout"] = {v: k for k, v in LAYOUTS.items()}.get(i.layout, str(i.layout))
        return d

    def close(self) -> None:
        if self._h:
            self._lib.kw_fd1d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Fd1dGpu_BlackScholes_Pricer(Fd1dGpu_Pricer):
    """Drop-in for kw::Fd1d_BlackScholes_Pricer (src/Pricer/kwFd1d_BlackScholes.cpp:15-43)."""

    _mode_bs = True


class PricerFactory:
    """kw::PricerFactory::create (src/Pricer/kwPricerFactory.h:15-41) with the GPU keys added."""

    MODES = {"FD1D-GPU": Fd1dGpu_Pricer, "FD1D-BS-GPU": Fd1dGpu_BlackScholes_Pricer}

    @staticmethod
    def create(config: Config) -> Tuple[str, Optional[Pricer]]:
        mode = config.get("PRICER", "")
        if mode == "":
            return "PricerFactory: Missing PRICER key", None
        cls = PricerFactory.MODES.get(mode)
        if cls is None:
            return "PricerFactory: Unknown PRICER = " + mode, None
        pricer = cls()
        err = pricer.init(config)
        if err:
            return "PricerFactory: " + err, None
        return "", pricer


def has_variant(variant: int, precision: str = "f64") -> bool:
    """True if this build of the library carries kernel variant `variant` (experiments need make EXPERIMENTS=1)."""
    L = load_library()
    return bool(L.kw_fd1d_has_variant(int(variant), PRECISIONS.get(precision, 0)))


def fp64_peak(device: int = 0) -> Tuple[float, float]:
    """Measured DFMA throughput (TFLOP/s) and the SM clock it implies at 64 DFMA/clk/SM."""
    L = load_library()
    t, mhz = C.c_double(0), C.c_double(0)
    rc = L.kw_fd1d_fp64_peak(device, C.byref(t), C.byref(mhz))
    if rc != KW_FD1D_OK:
        raise RuntimeError("kw_fd1d_fp64_peak failed (no CUDA device?)")
    return t.value, mhz.value


def microbench(device: int = 0) -> dict:
    L = load_library()
    out = (C.c_double * 8)()
    rc = L.kw_fd1d_microbench(device, out)
    if rc != KW_FD1D_OK:
        raise RuntimeError("kw_fd1d_microbench failed (no CUDA device?)")
    names = ["dfma_dep", "shfl64_dep", "shfl64_dfma_dep", "syncthreads_4warps", "lds_dep", "dadd_dmnmx_dep"]
    return {k: out[i] for i, k in enumerate(names)}


def tmem_probe(device: int = 0) -> dict:
    """Cycles per round (64 doubles per thread from TMEM, 16 warps per SM) of the tensor-memory probes."""
    L = load_library()
    out = (C.c_double * 16)()
    rc = L.kw_fd1d_tmem_probe(device, out)
    if rc != KW_FD1D_OK:
        raise RuntimeError("kw_fd1d_tmem_probe failed (no CUDA device?)")
    names = ["dfma16_only", "ld_only", "ld_dfma8", "ld_dfma16", "ld_only_b", "ld_ahead_dfma16", "one_warp_ld",
             "mismatches"]
    res = {k: out[i] for i, k in enumerate(names)}
    res["coresident_ctas"] = [int(out[8 + i]) for i in range(6)]
    res["occupancy"] = int(out[14])
    res["clock64_ticks_per_ns"] = out[15]
    res["tmem_bytes_per_clk_sm"] = res["coresident_ctas"][1] * 128 * 64 * 8 / out[1] if out[1] else None
    return res


def dfma_probe(device: int = 0) -> dict:
    """DFMA TFLOP/s by number of distinct register source operands, at 16 and 64 warps per SM."""
    L = load_library()
    out = (C.c_double * 10)()
    if L.kw_fd1d_dfma_probe(device, out) != KW_FD1D_OK:
        raise RuntimeError("kw_fd1d_dfma_probe failed (no CUDA device?)")
    names = ["1reg", "2reg", "3reg", "3reg_shared"]
    res = {f"{n}_{w}warps": out[i * 4 + j] for i, w in enumerate((16, 64)) for j, n in enumerate(names)}
    res["3reg_plus_1sel_16warps"] = out[8]
    res["3reg_plus_2sel_16warps"] = out[9]
    return res
