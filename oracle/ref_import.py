"""TEST INFRASTRUCTURE ONLY -- import the reference's own modules, unmodified, by path.

Works only where ``/root/reference`` exists (the build container).  The GPU box has no
reference tree; everything that must run there uses :mod:`oracle.hotpath` and the committed
fixtures under ``tests/golden/`` instead.  Uses the same ``importlib`` path-loading technique
the reference itself uses (10_benchmark_folder_detection.py:80-95).
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
from pathlib import Path
from types import ModuleType

from . import shims

REFERENCE_ROOT = Path(os.environ.get("AVLD_REFERENCE_ROOT", "/root/reference"))
LSE = REFERENCE_ROOT / "latent_space_exploration"


def available() -> bool:
    return (LSE / "map_detector_core.py").exists()


def _load_by_path(name: str, path: Path) -> ModuleType:
    spec = importlib.util.spec_from_file_location(name, str(path))
    if spec is None or spec.loader is None:
        raise ImportError(f"cannot load {path}")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_cache: dict = {}


def load(which: str) -> ModuleType:
    """``which`` in {'00', '07', '08', '09', '10', 'core'} -> the reference module object."""
    if which in _cache:
        return _cache[which]
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    shims.install()
    files = {
        "00": "00_normalize_dataset_rms.py",
        "07": "07_encode_wav_to_latent.py",
        "08": "08_fit_radial_detector.py",
        "09": "09_evaluate_wav_detection.py",
    }
    if which == "core":
        if str(REFERENCE_ROOT) not in sys.path:
            sys.path.insert(0, str(REFERENCE_ROOT))
        # our repo also ships a ``latent_space_exploration`` package (the drop-in surface);
        # load the reference's core under a private name so the two never alias.
        mod = _load_by_path("_ref_map_detector_core", LSE / "map_detector_core.py")
    elif which == "10":
        try:
            import matplotlib  # noqa: F401
        except Exception:
            _install_matplotlib_stub()
        mod = _load_by_path("_ref_10_benchmark", LSE / "10_benchmark_folder_detection.py")
    elif which in files:
        mod = _load_by_path(f"_ref_{which}", LSE / files[which])
    else:
        raise KeyError(which)
    _cache[which] = mod
    return mod


def _install_matplotlib_stub() -> None:
    """10_benchmark_folder_detection.py imports matplotlib for its (out-of-scope) plots."""
    import types

    class _Anything(types.ModuleType):
        """Every attribute is a callable that returns another such object (plots are out of scope)."""

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Anything(name)

        def __call__(self, *a, **k):
            return _Anything("result")

        def __iter__(self):
            return iter((_Anything("a"), _Anything("b")))

    mpl = _Anything("matplotlib")
    plt = _Anything("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
