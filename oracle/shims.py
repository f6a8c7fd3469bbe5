"""TEST INFRASTRUCTURE ONLY -- ``sys.modules`` stand-ins that let the unmodified reference import.

Every hot-path module of the reference does a module-level ``import librosa`` that raises
``SystemExit`` when missing (map_detector_core.py:32-35, 08_fit_radial_detector.py:52-55 ...).
``install()`` registers minimal modules named ``librosa``, ``soundfile``, ``omegaconf`` and
``hydra.utils`` that forward to :mod:`oracle.librosa_port` / PyYAML, so the reference's own
control flow (crop/pad, z-score, encoder call, output unpacking, fit, decision) executes
verbatim.  Nothing is installed if the real package is importable.
"""
from __future__ import annotations

import importlib
import sys
import types

from . import librosa_port as _lp


def _have(name: str) -> bool:
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


def _make_librosa() -> types.ModuleType:
    m = types.ModuleType("librosa")
    m.__version__ = "0.9.2+oracle-port"
    m.load = _lp.load
    m.stft = _lp.stft
    m.power_to_db = _lp.power_to_db
    feature = types.ModuleType("librosa.feature")
    feature.melspectrogram = _lp.melspectrogram
    filters = types.ModuleType("librosa.filters")
    filters.mel = _lp.mel_filterbank
    m.feature = feature
    m.filters = filters
    sys.modules["librosa.feature"] = feature
    sys.modules["librosa.filters"] = filters
    return m


def _make_soundfile() -> types.ModuleType:
    m = types.ModuleType("soundfile")
    m.__version__ = "0.13.1+oracle-port"

    def write(file, data, samplerate, subtype=None, **_):
        _lp.write_wav(file, data, samplerate, subtype=subtype)

    def read(file, dtype="float32", always_2d=False, **_):
        x, sr = _lp.read_wav(file)
        if not always_2d and x.shape[1] == 1:
            x = x[:, 0]
        return x.astype(dtype), sr

    m.write = write
    m.read = read
    return m


def _make_omegaconf_and_hydra():
    import yaml

    class OmegaConf:  # the two calls the reference makes (map_detector_core.py:110-111)
        @staticmethod
        def load(path):
            with open(path, "r", encoding="utf-8") as f:
                return yaml.safe_load(f)

        @staticmethod
        def to_container(cfg, resolve=False):
            return cfg

    def instantiate(cfg, *args, **kwargs):
        """``hydra.utils.instantiate`` for a flat ``{_target_: 'pkg.mod.Class', **kw}`` node."""
        cfg = dict(cfg)
        target = cfg.pop("_target_")
        cfg.pop("_partial_", None)
        mod_name, _, attr = target.rpartition(".")
        obj = getattr(importlib.import_module(mod_name), attr)
        cfg.update(kwargs)
        return obj(*args, **cfg)

    oc = types.ModuleType("omegaconf")
    oc.OmegaConf = OmegaConf
    hydra = types.ModuleType("hydra")
    hutils = types.ModuleType("hydra.utils")
    hutils.instantiate = instantiate
    hydra.utils = hutils
    return oc, hydra, hutils


def install() -> dict:
    """Install the shims that are needed; returns ``{name: 'real'|'shim'}``."""
    status = {}
    if _have("librosa") and not getattr(sys.modules.get("librosa"), "__version__", "").endswith("oracle-port"):
        status["librosa"] = "real"
    else:
        sys.modules["librosa"] = _make_librosa()
        status["librosa"] = "shim"
    if _have("soundfile") and not getattr(sys.modules.get("soundfile"), "__version__", "").endswith("oracle-port"):
        status["soundfile"] = "real"
    else:
        sys.modules["soundfile"] = _make_soundfile()
        status["soundfile"] = "shim"
    if _have("omegaconf") and _have("hydra.utils") and not hasattr(sys.modules["omegaconf"], "_oracle_shim"):
        status["omegaconf/hydra"] = "real"
    else:
        oc, hydra, hutils = _make_omegaconf_and_hydra()
        oc._oracle_shim = True
        sys.modules["omegaconf"] = oc
        sys.modules["hydra"] = hydra
        sys.modules["hydra.utils"] = hutils
        status["omegaconf/hydra"] = "shim"
    return status
