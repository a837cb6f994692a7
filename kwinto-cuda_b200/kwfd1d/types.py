"""Wire types of the Fd1d path.

``OPTION_DTYPE`` is layout-identical to the reference's ``kw::Option``
(src/Core/kwAsset.h:12-22: f64 t,k,z,r,q,s; u8 e; i8 w -- 56 bytes, offsets
0/8/16/24/32/40/48/49) and to ``kw_option`` in include/kw_fd1d.h, so a numpy array of this
dtype, a ``std::vector<kw::Option>::data()`` and the C-ABI all see the same bytes.
"""
import numpy as np

OPTION_DTYPE = np.dtype(
    {
        "names": ["t", "k", "z", "r", "q", "s", "e", "w"],
        "formats": ["<f8", "<f8", "<f8", "<f8", "<f8", "<f8", "u1", "i1"],
        "offsets": [0, 8, 16, 24, 32, 40, 48, 49],
        "itemsize": 56,
    }
)


def make_options(rows) -> np.ndarray:
    """rows of (t, k, z, r, q, s, e, w) -- the reference's aggregate-init order
    (test/kwPricer_test.cpp:16)."""
    return np.array([tuple(r) for r in rows], dtype=OPTION_DTYPE)
