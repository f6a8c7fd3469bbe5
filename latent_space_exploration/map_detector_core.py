"""Module surface of the reference's ``map_detector_core`` (imported by 08b / 09n / 10b as
``from latent_space_exploration.map_detector_core import ...``, 08b:45-57): the same names, implemented on the GPU library."""
from amphibian_vae_latent_detector_b200.reference_api import (  # noqa: F401
    build_nn_module, crop_or_pad_time, encode_wav_to_latent, gaussian_logpdf_from_precision, get_chunk_seconds_for_map,
    get_priors_from_map_meta, inv_and_logdet, load_encoder, load_json, load_yaml_cfg, pick_encoder_cfg, read_map_detector_params,
    split_model_and_state, wav_to_mel)
from amphibian_vae_latent_detector_b200.cli import find_project_root  # noqa: F401,E402

import json as _json  # noqa: E402
from pathlib import Path as _Path  # noqa: E402

import numpy as _np  # noqa: E402

_ENC_DIR = ("models", "bird_net_vae_audio_splitted_encoder_v0")               # core:64-77 looks under models/, not downloaded_models/


def _must_exist(p: _Path, what: str) -> _Path:
    if not p.exists():
        raise FileNotFoundError(f"No encontré {what} en: {p}")
    return p


def resolve_default_config(project_root: _Path) -> _Path:
    return _must_exist(_Path(project_root) / "config.json", "config.json")


def resolve_default_encoder_pt(project_root: _Path) -> _Path:
    return _must_exist(_Path(project_root).joinpath(*_ENC_DIR, "model.pt"), "encoder .pt")


def resolve_default_encoder_yaml(project_root: _Path) -> _Path:
    return _must_exist(_Path(project_root).joinpath(*_ENC_DIR, "bird_net_vae_audio_splitted.yaml"), "encoder YAML")


def save_json(path, obj) -> None:
    _Path(path).write_text(_json.dumps(obj, indent=2, ensure_ascii=False), encoding="utf-8")


def summarize_1d(x) -> dict:
    """min / p05 / p50 / p95 / max of a 1-D array, NaN for an empty one (core:92-101)."""
    x = _np.asarray(x)
    keys = ("min", "p05", "p50", "p95", "max")
    if x.size == 0:
        return dict.fromkeys(keys, float("nan"))
    return dict(zip(keys, (float(v) for v in _np.quantile(x, [0.0, 0.05, 0.5, 0.95, 1.0]))))
