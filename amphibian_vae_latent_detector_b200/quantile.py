"""Host-side interpolation of ``np.quantile(x, q)`` (method="linear") from exact order statistics.

The GPU selects the two neighbouring order statistics exactly (``avld_order_stats``); this module
decides *which* ranks are needed and blends them the way numpy does, so that the thresholds written
to ``config.json`` match ``quantile_safe`` (08_fit_radial_detector.py:109-112) bit for bit:

* ``semantics="numpy2"`` (numpy >= 2, what the in-container oracle executes): a Python-float ``q`` is
  cast to the array dtype, so for float32 radii the virtual index ``(n-1)*q`` and the lerp are
  evaluated in float32.
* ``semantics="numpy1"`` (numpy 1.26.4, the reference's pin): ``q`` stays float64; the index and the
  blend are float64.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def neighbour_ranks(n: int, q: float, semantics: str = "numpy2") -> Tuple[int, int, float]:
    """-> (previous_rank, next_rank, gamma) for a sorted float32 array of length ``n``."""
    if n <= 0:
        raise ValueError("empty population")
    if not (0.0 <= q <= 1.0):
        raise ValueError("Quantiles must be in the range [0, 1]")
    if semantics == "numpy2":
        qq = np.float32(q)
        vi = np.float32(n - 1) * qq                       # (n - 1) * quantiles, float32
    elif semantics == "numpy1":
        vi = np.float64(n - 1) * np.float64(q)
    else:
        raise ValueError(semantics)
    prev = int(np.floor(vi))
    nxt = prev + 1
    if vi >= n - 1:
        prev = nxt = n - 1
    if vi < 0:
        prev = nxt = 0
    gamma = float(vi) - float(prev)                       # exact; cast back to the index dtype below
    return prev, nxt, gamma


def lerp(a: float, b: float, gamma: float, semantics: str = "numpy2") -> float:
    """numpy ``_lerp``: ``a + (b-a)*t``, replaced by ``b - (b-a)*(1-t)`` where ``t >= 0.5``."""
    a32, b32 = np.float32(a), np.float32(b)
    diff = np.float32(b32 - a32)
    if semantics == "numpy2":
        t = np.float32(gamma)
        out = np.float32(a32 + np.float32(diff * t))
        if t >= 0.5:
            out = np.float32(b32 - np.float32(diff * np.float32(np.float32(1) - t)))
        return float(out)
    t = np.float64(gamma)
    out = np.float64(a32) + np.float64(diff) * t
    if t >= 0.5:
        out = np.float64(b32) - np.float64(diff) * (1 - t)
    return float(out)
