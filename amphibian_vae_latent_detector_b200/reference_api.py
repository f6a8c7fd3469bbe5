"""The reference's own function surface for the hot path -- same names, arguments, return types and
error behaviour -- executed by the CUDA library.  Paths relative to
``latent_space_exploration/`` in the reference:

    rms_normalize, process_folder                       00_normalize_dataset_rms.py:29-57
    crop_or_pad_time, wav_to_mel, encode_wav_to_latent  map_detector_core.py:185-300 (= 07/08/09 copies)
    load_encoder (+ helpers)                            map_detector_core.py:104-179
    l2_norm_rows, quantile_safe, summarize_dist,
    fit_species_with_fp_control                         08_fit_radial_detector.py:105-123, :310-333
    get_detector_from_config, l2, detect_species,
    PRIORITY_ORDER                                      09_evaluate_wav_detection.py:61-66, :113-149, :354-436
    DetectorSession                                     10_benchmark_folder_detection.py:113-199

plus batched additions (``*_batch`` / ``predict_many``) that feed many files per GPU pass.  File I/O is
WAV via the standard library (the reference's librosa.load on PCM_16 / 24 / 32 / U8 / IEEE-float files of any
channel count, soundfile.write as PCM_16); resampling is not implemented (the reference's datasets are already 48 kHz) and raises.
Everything numeric runs on the GPU: ``device`` arguments are accepted for signature compatibility, and
"cpu" is mapped to ``cuda:0`` -- there is no CPU implementation to fall back to.
"""
from __future__ import annotations

import os

import importlib
import json
import sys
import wave
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple

from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from .engine import Engine, priority_ranks

PRIORITY_ORDER = [
    "Batrachyla_leptopus",
    "Batrachyla_taeniata",
    "Calyptocephalella_gayi",
    "Pleurodema_thaul",
]

# Engine cache.  An engine owns a device context (operand scratch of ~3 MB per chunk of `max_batch`, plus the uploaded layer
# program), so the cache is a small LRU (an evicted engine is freed as soon as nobody else holds it).  Entries with an encoder keep a strong reference to the
# module (an id() alone could be reused by a new module after garbage collection) and are keyed by a fingerprint of its
# weights as well (an in-place load_state_dict on the same object must not be served the old program).
_ENGINE_CACHE_SIZE = 4
_ENGINES: "OrderedDict[tuple, tuple]" = OrderedDict()     # key -> (engine, encoder or None)
_IO_POOL = ThreadPoolExecutor(max_workers=max(2, min(16, os.cpu_count() or 4)))
_LOADED: Dict[tuple, int] = {}


def _cuda_index(device) -> int:
    dev = torch.device(device) if not isinstance(device, torch.device) else device
    return (dev.index or 0) if dev.type == "cuda" else 0


def _cache_get(key):
    hit = _ENGINES.get(key)
    if hit is None:
        return None
    _ENGINES.move_to_end(key)
    return hit[0]


def _cache_put(key, eng: Engine, encoder=None) -> Engine:
    _ENGINES[key] = (eng, encoder)
    while len(_ENGINES) > _ENGINE_CACHE_SIZE:
        _ENGINES.popitem(last=False)      # dropped, not closed: a caller may still hold it; Engine.__del__ frees the context
    return eng


def clear_engine_cache() -> None:
    """Close every cached engine (frees their device memory)."""
    while _ENGINES:
        _, (old, _enc) = _ENGINES.popitem()
        old.close()


def _weights_fingerprint(encoder) -> tuple:
    """Cheap identity of the module's current weights: (number of tensors, total elements, a float64 checksum of strided
    samples of every parameter / buffer).  Two state dicts that differ anywhere a sample falls differ here; an in-place
    update of the weights changes it with overwhelming probability."""
    n, total, acc = 0, 0, 0.0
    with torch.no_grad():
        for t in list(encoder.parameters()) + list(encoder.buffers()):
            flat = t.detach().reshape(-1)
            if flat.numel() == 0 or not (flat.is_floating_point() or flat.dtype in (torch.int64, torch.int32)):
                continue
            step = max(1, flat.numel() // 64)
            acc += float(flat[::step].double().sum().item()) * (n + 1)
            n += 1
            total += flat.numel()
    return (n, total, acc)


def _engine(chunk_len: int, device=0, *, sr=48000, n_mels=64, fmin=150.0, fmax=15000.0, hop_length=384, n_fft=2048,
            target_frames=192, max_batch: int = 64) -> Engine:
    key = (_cuda_index(device), int(chunk_len), sr, n_mels, float(fmin), float(fmax), hop_length, n_fft, target_frames)
    eng = _cache_get(key)
    if eng is None:
        eng = _cache_put(key, Engine(key[0], chunk_len=int(chunk_len), max_batch=max_batch, sr=sr, n_fft=n_fft,
                                     hop_length=hop_length, n_mels=n_mels, fmin=fmin, fmax=fmax, target_frames=target_frames))
    return eng


def _engine_with_encoder(encoder, chunk_len, device, *, max_batch: int = 64, **mel_kw) -> Engine:
    """One engine per (geometry, encoder object + weights, pass size); the layer program is exported and uploaded once."""
    key = (_cuda_index(device), int(chunk_len), tuple(sorted(mel_kw.items())), id(encoder), _weights_fingerprint(encoder),
           int(max_batch))
    eng = _cache_get(key)
    if eng is None:
        eng = Engine(key[0], chunk_len=int(chunk_len), max_batch=int(max_batch), **mel_kw)
        eng.load_encoder(encoder)
        _cache_put(key, eng, encoder)
    return eng


# ----------------------------------------------------------------------------------------------------------
# WAV I/O (librosa.load(sr=sr, mono=True) on PCM files; soundfile.write(float data) = PCM_16)
# ----------------------------------------------------------------------------------------------------------
def _read_ieee_float_wav(path):
    """RIFF/WAVE with format tag 3 (IEEE float; also the EXTENSIBLE form), which the standard library's ``wave`` rejects:
    -> (channels, bytes per sample, rate, raw payload)."""
    import struct
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            raise RuntimeError(f"{path}: not a RIFF/WAVE file")
        fmt = None
        while True:
            hdr = f.read(8)
            if len(hdr) < 8:
                raise RuntimeError(f"{path}: no data chunk")
            name, size = hdr[:4], struct.unpack("<I", hdr[4:])[0]
            if name == b"fmt ":
                body = f.read(size + (size & 1))
                fmt = struct.unpack("<HHIIHH", body[:16])
                if fmt[0] == 0xFFFE and len(body) >= 26:                          # EXTENSIBLE: the real tag leads the GUID
                    fmt = (struct.unpack("<H", body[24:26])[0],) + fmt[1:]
            elif name == b"data":
                raw = f.read(size)
                break
            else:
                f.seek(size + (size & 1), 1)
    if fmt is None or fmt[0] != 3 or fmt[5] not in (32, 64):
        raise RuntimeError(f"{path}: unsupported WAV format tag {None if fmt is None else fmt[0]}")
    return fmt[1], fmt[5] // 8, fmt[2], raw


def load_wav(path, sr: int = 48000) -> np.ndarray:
    """``librosa.load(path, sr=sr, mono=True)`` for WAV files: float32 samples exactly as libsndfile converts them
    (PCM_16 / 2^15, PCM_24 / 2^23, PCM_32 / 2^31, PCM_U8 (u - 128) / 2^7, IEEE float as stored), channels averaged.
    A file whose rate differs from ``sr`` is resampled as librosa 0.9.2 does (``res_type="kaiser_best"``) -- on the GPU
    (``avld_resample``), after the mono mix, like ``librosa.load``."""
    is_float = False
    try:
        with wave.open(str(path), "rb") as w:
            nch, width, rate, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
            raw = w.readframes(n)
    except wave.Error as e:
        if "unknown format" not in str(e) and "unknown extended format" not in str(e):
            raise
        nch, width, rate, raw = _read_ieee_float_wav(path)
        is_float = True
    if is_float:
        x = np.frombuffer(raw, dtype="<f4" if width == 4 else "<f8").astype(np.float32)
    elif width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) * np.float32(1.0 / 32768.0)
    elif width == 4:
        x = (np.frombuffer(raw, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = (b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16))
        v = np.where(v >= 1 << 23, v - (1 << 24), v)                               # sign-extend 24 -> 32 bits
        x = v.astype(np.float32) * np.float32(1.0 / 8388608.0)                     # exact: 24-bit integers fit a float32
    elif width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - np.float32(128.0)) * np.float32(1.0 / 128.0)
    else:
        raise RuntimeError(f"{path}: unsupported sample width {width}")
    if nch > 1:
        x = x.reshape(-1, nch).mean(axis=1).astype(np.float32)     # librosa to_mono
    x = np.ascontiguousarray(x)
    if sr is not None and rate != sr:                              # librosa.load: resample after to_mono
        eng = _engine(144000, 0)
        x = eng.resample(torch.from_numpy(x).to(eng.device), int(rate), int(sr)).cpu().numpy()
    return x


def write_wav_pcm16(path, y: np.ndarray, sr: int) -> None:
    """``sf.write(path, y, sr)`` for float data on a .wav path: PCM_16, ``lrintf(x * 0x7FFF)``."""
    pcm = np.clip(np.rint(np.asarray(y, dtype=np.float32) * np.float32(32767.0)), -32768, 32767).astype("<i2")
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(sr))
        w.writeframes(pcm.tobytes())


def _fix_length(y: np.ndarray, sr: int, duration: float) -> np.ndarray:
    if duration > 0:                                               # core:212-217
        target_len = int(sr * duration)
        if y.shape[0] < target_len:
            y = np.pad(y, (0, target_len - y.shape[0]), mode="constant")
        else:
            y = y[:target_len]
    return y


# ----------------------------------------------------------------------------------------------------------
# 00_normalize_dataset_rms.py
# ----------------------------------------------------------------------------------------------------------
def rms_normalize(y, target_rms=0.05, rms_min=1e-4, eps=1e-8):
    """00:29-38.  -> ``(y_norm float32 ndarray, ok bool)``; silent input is returned unchanged."""
    y = np.ascontiguousarray(y, dtype=np.float32)
    if y.ndim != 1 or y.size == 0:
        raise ValueError("rms_normalize expects a non-empty 1-D signal")
    eng = _engine(y.shape[0], 0, max_batch=8)
    out, ok, _ = eng.rms_normalize(torch.from_numpy(y[None]).to(eng.device), target_rms, rms_min, eps)
    return out[0].cpu().numpy(), bool(ok[0].item())


def rms_normalize_batch(x: np.ndarray, target_rms=0.05, rms_min=1e-4, eps=1e-8):
    """rows of ``x [n, L]`` -> ``(y [n, L], ok [n] bool)``."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    eng = _engine(x.shape[1], 0, max_batch=64)
    out, ok, _ = eng.rms_normalize(torch.from_numpy(x).to(eng.device), target_rms, rms_min, eps)
    return out.cpu().numpy(), ok.cpu().numpy().astype(bool)


def process_folder(src_root: Path, dst_root: Path, sr=48000):
    """00:41-57: every ``<species>/*.wav`` -> normalised PCM_16 WAV under ``dst_root`` (written even when
    the file is silent).  Files of equal length are normalised in one GPU batch."""
    src_root, dst_root = Path(src_root), Path(dst_root)
    for species_dir in sorted(src_root.iterdir()):
        if not species_dir.is_dir():
            continue
        out_dir = dst_root / species_dir.name
        out_dir.mkdir(parents=True, exist_ok=True)
        by_len: Dict[int, List[Tuple[Path, np.ndarray]]] = {}
        for wav in sorted(species_dir.glob("*.wav")):
            y = load_wav(wav, sr)
            by_len.setdefault(y.shape[0], []).append((wav, y))
        for length, items in by_len.items():
            ys, _ = rms_normalize_batch(np.stack([y for _, y in items]))
            for (wav, _), yn in zip(items, ys):
                write_wav_pcm16(out_dir / wav.name, yn, sr)


# ----------------------------------------------------------------------------------------------------------
# map_detector_core.py: features and encoding
# ----------------------------------------------------------------------------------------------------------
def crop_or_pad_time(mel: np.ndarray, target_frames: int) -> np.ndarray:
    """core:185-195 (host-side helper kept for API compatibility; the GPU path crops inside the kernel)."""
    _, T = mel.shape
    if T == target_frames:
        return mel
    if T > target_frames:
        start = (T - target_frames) // 2
        return mel[:, start:start + target_frames]
    pad_total = target_frames - T
    pad_left = pad_total // 2
    return np.pad(mel, ((0, 0), (pad_left, pad_total - pad_left)), mode="constant")


def wav_to_mel(wav_path: Path, *, sr: int = 48000, duration: float = 5.0, n_mels: int = 64, fmin: float = 150.0,
               fmax: float = 15000.0, hop_length: int = 384, n_fft: int = 2048, target_frames: int = 192) -> torch.Tensor:
    """core:198-237 -> ``torch.float32 [n_mels, target_frames]`` (on the CPU, like the reference)."""
    y = _fix_length(load_wav(wav_path, sr), sr, duration)
    eng = _engine(y.shape[0], 0, sr=sr, n_mels=n_mels, fmin=fmin, fmax=fmax, hop_length=hop_length, n_fft=n_fft,
                  target_frames=target_frames)
    feat = eng.logmel(torch.from_numpy(y[None]).to(eng.device))[0]              # [T, M]
    return feat.T.contiguous().cpu()


def encode_wavs_to_latents(encoder: torch.nn.Module, wav_paths: Sequence[Path], device=None, *, sr: int = 48000,
                           duration: float = 5.0, n_mels: int = 64, fmin: float = 150.0, fmax: float = 15000.0,
                           hop_length: int = 384, n_fft: int = 2048, target_frames: int = 192,
                           return_failed: bool = False):
    """Batched ``encode_wav_to_latent``: many files per GPU pass.  Unreadable files are skipped and counted
    (the reference's per-file try/except, 08:489-506)."""
    mel_kw = dict(sr=sr, n_mels=n_mels, fmin=fmin, fmax=fmax, hop_length=hop_length, n_fft=n_fft,
                  target_frames=target_frames)
    chunk_len = int(sr * duration)
    # decode on a few threads (file reads and numpy conversions release the GIL), slabs of files so that a large folder
    # never sits in host memory at once; order and the count-and-continue semantics are those of the per-file loop
    def _load(p):
        try:
            return _fix_length(load_wav(p, sr), sr, duration)
        except Exception:
            return None

    wav_paths = list(wav_paths)
    eng = _engine_with_encoder(encoder, chunk_len, device if device is not None else 0,
                               max_batch=64 if len(wav_paths) <= 256 else 512, **mel_kw)
    failed, parts = [], []
    for i in range(0, len(wav_paths), 2048):
        piece = wav_paths[i:i + 2048]
        rows = list(_IO_POOL.map(_load, piece)) if len(piece) > 8 else [_load(p) for p in piece]
        failed += [p for p, r in zip(piece, rows) if r is None]
        rows = [r for r in rows if r is not None]
        if rows:
            x = torch.from_numpy(np.stack(rows)).to(eng.device)
            feat = eng.logmel(x)                   # files on disk are already normalised (00) -> no RMS stage here
            parts.append(eng.encoder_forward(feat).cpu().numpy())
    Z = np.concatenate(parts) if parts else np.zeros((0, eng.latent_dim), dtype=np.float32)
    return (Z, failed) if return_failed else Z


@torch.no_grad()
def encode_wav_to_latent(encoder: torch.nn.Module, wav_path: Path, device=None, *, sr: int = 48000,
                         duration: float = 5.0, n_mels: int = 64, fmin: float = 150.0, fmax: float = 15000.0,
                         hop_length: int = 384, n_fft: int = 2048, target_frames: int = 192) -> np.ndarray:
    """core:241-300 -> ``np.float32 [D]``."""
    Z = encode_wavs_to_latents(encoder, [wav_path], device, sr=sr, duration=duration, n_mels=n_mels, fmin=fmin,
                               fmax=fmax, hop_length=hop_length, n_fft=n_fft, target_frames=target_frames,
                               return_failed=True)
    if Z[1]:
        load_wav(wav_path, sr)                     # re-raise the file's own error, as the reference would
    return Z[0][0].astype(np.float32)


# ---- encoder loading: host-only logic, restated from core:104-179 -------------------------------------------
def load_yaml_cfg(cfg_path: Path) -> Dict[str, Any]:
    import yaml
    with open(cfg_path, "r", encoding="utf-8") as f:
        cfg = yaml.safe_load(f)                    # == OmegaConf.to_container(resolve=False) for plain YAML
    if not isinstance(cfg, dict):
        raise ValueError("YAML no es dict tras to_container.")
    return cfg


def pick_encoder_cfg(cfg: Dict[str, Any]) -> Dict[str, Any]:
    enc = cfg.get("encoder")
    if not isinstance(enc, dict) or "_target_" not in enc:
        raise KeyError("No pude encontrar cfg['encoder'] con _target_ en el YAML.")
    return enc


def split_model_and_state(ckpt: Any):
    if isinstance(ckpt, torch.nn.Module):
        return ckpt, None
    if isinstance(ckpt, dict):
        for key in ("state_dict", "model_state_dict"):
            if key in ckpt and isinstance(ckpt[key], dict):
                return None, ckpt[key]
        if ckpt and all(isinstance(v, torch.Tensor) for v in ckpt.values()):
            return None, ckpt
    raise RuntimeError("Checkpoint no reconocido (ni nn.Module ni state_dict).")


def build_nn_module(obj: Any) -> torch.nn.Module:
    if isinstance(obj, torch.nn.Module):
        return obj
    if callable(obj):
        out = obj()
        if isinstance(out, torch.nn.Module):
            return out
    raise RuntimeError(f"instantiate devolvió {type(obj)}; no pude obtener nn.Module.")


def _instantiate(cfg):
    """The part of ``hydra.utils.instantiate`` the encoder YAMLs use (core:171): ``_target_`` (dotted path of a callable),
    nested ``_target_`` nodes in arguments / lists are instantiated first, ``_partial_: true`` returns a
    ``functools.partial`` instead of calling, ``_args_`` gives positional arguments; other keys are keyword arguments."""
    if isinstance(cfg, (list, tuple)):
        return [_instantiate(v) for v in cfg]
    if not isinstance(cfg, dict):
        return cfg
    if "_target_" not in cfg:
        return {k: _instantiate(v) for k, v in cfg.items()}
    cfg = dict(cfg)
    target = cfg.pop("_target_")
    partial = bool(cfg.pop("_partial_", False))
    args = [_instantiate(v) for v in cfg.pop("_args_", [])]
    cfg.pop("_convert_", None)
    cfg.pop("_recursive_", None)
    kwargs = {k: _instantiate(v) for k, v in cfg.items()}
    if callable(target):
        fn = target
    else:
        mod_name, _, attr = str(target).rpartition(".")
        fn = getattr(importlib.import_module(mod_name), attr)
    if partial:
        import functools
        return functools.partial(fn, *args, **kwargs)
    return fn(*args, **kwargs)


def load_encoder(encoder_pt: Path, encoder_yaml: Path, project_root: Path, device=None, *,
                 trust_checkpoint: Optional[bool] = None) -> torch.nn.Module:
    """core:150-179.  Returns the ``nn.Module`` (eval mode, on the CPU: it is only walked by ``export_program``; the forward
    itself runs in the CUDA library).

    The reference calls ``torch.load`` with torch 2.1's default (full unpickling) and so also accepts a checkpoint that holds
    the pickled module itself (core:160-165).  Here a checkpoint is first read with ``weights_only=True`` (tensors and plain
    containers: every state-dict checkpoint); one that needs arbitrary unpickling is only loaded when the caller says the file
    is trusted -- ``trust_checkpoint=True`` or ``AVLD_TRUST_CHECKPOINTS=1`` -- because unpickling executes code from the file."""
    if str(project_root) not in sys.path:
        sys.path.insert(0, str(project_root))
    encoder_pt = Path(encoder_pt)
    if not encoder_pt.exists():
        raise FileNotFoundError(f"No existe encoder .pt: {encoder_pt}")
    try:
        ckpt = torch.load(str(encoder_pt), map_location="cpu", weights_only=True)
    except Exception as exc:                       # pickle.UnpicklingError and friends: not a plain state dict
        if trust_checkpoint is None:
            trust_checkpoint = os.environ.get("AVLD_TRUST_CHECKPOINTS", "0") not in ("", "0")
        if not trust_checkpoint:
            raise RuntimeError(
                f"{encoder_pt} is not a plain state-dict checkpoint (torch.load(weights_only=True) refused it: {exc}).  "
                "Full-module checkpoints execute code when unpickled; pass trust_checkpoint=True (or set "
                "AVLD_TRUST_CHECKPOINTS=1) if the file comes from a trusted source.") from exc
        ckpt = torch.load(str(encoder_pt), map_location="cpu", weights_only=False)
    model, state = split_model_and_state(ckpt)
    if model is not None:
        return build_nn_module(model).eval()
    if state is None:
        raise RuntimeError("No encontré state_dict en el checkpoint.")
    encoder_yaml = Path(encoder_yaml)
    if not encoder_yaml.exists():
        raise FileNotFoundError(f"No existe YAML: {encoder_yaml}")
    model = build_nn_module(_instantiate(pick_encoder_cfg(load_yaml_cfg(encoder_yaml))))
    model.load_state_dict(state, strict=False)
    return model.eval()


# ----------------------------------------------------------------------------------------------------------
# 08_fit_radial_detector.py
# ----------------------------------------------------------------------------------------------------------
def _dev_rows(x: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to("cuda:0")


def l2_norm_rows(x: np.ndarray) -> np.ndarray:
    """08:105-106 via the radii kernel against a zero centroid."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.shape[0] == 0:
        return np.zeros(0, dtype=np.float32)
    eng = _engine(144000, 0)
    zero = torch.zeros(1, x.shape[1], dtype=torch.float32, device=eng.device)
    return eng.radii(_dev_rows(x), zero)[:, 0].cpu().numpy()


def _quantiles(x: np.ndarray, qs: Sequence[float]) -> List[float]:
    from . import quantile as _q
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    eng = _engine(144000, 0)
    r = _dev_rows(x[:, None])
    lab = torch.zeros(x.shape[0], dtype=torch.int32, device=eng.device)
    queries, plan = [], []
    for q in qs:
        prev, nxt, gamma = _q.neighbour_ranks(x.shape[0], q)
        queries += [(0, 0, prev), (0, 0, nxt)]
        plan.append(gamma)
    vals = eng.order_stats(r, lab, queries)
    return [_q.lerp(float(vals[2 * i]), float(vals[2 * i + 1]), g) for i, g in enumerate(plan)]


def quantile_safe(x: np.ndarray, q: float) -> float:
    """08:109-112: 0.0 on empty input, else ``float(np.quantile(x, q))`` (exact selection on the GPU)."""
    if np.asarray(x).size == 0:
        return 0.0
    return _quantiles(x, [q])[0]


def summarize_dist(x: np.ndarray) -> Dict[str, float]:
    """08:115-123."""
    if np.asarray(x).size == 0:
        return {"min": float("nan"), "p50": float("nan"), "p90": float("nan"), "max": float("nan")}
    v = _quantiles(x, [0.0, 0.5, 0.9, 1.0])
    return {"min": v[0], "p50": v[1], "p90": v[2], "max": v[3]}


def fit_species_with_fp_control(Z_in: np.ndarray, Z_out: Optional[np.ndarray], q_in: float, q_out: float):
    """08:310-333 -> ``(mu float32[D], rk, rk_in, rk_out, extra)``."""
    Z_in = np.ascontiguousarray(Z_in, dtype=np.float32)
    has_out = Z_out is not None and np.asarray(Z_out).size > 0
    Z = np.concatenate([Z_in, np.ascontiguousarray(Z_out, dtype=np.float32)]) if has_out else Z_in
    lab = np.zeros(Z.shape[0], dtype=np.int32)
    lab[Z_in.shape[0]:] = 1
    eng = _engine(144000, 0)
    fit = eng.fit_radial(_dev_rows(Z), torch.from_numpy(lab).to(eng.device), 2 if has_out else 1, q_in, q_out)
    names = ("min", "p50", "p90", "max")
    nan4 = {k: float("nan") for k in names}
    extra = {"rho_in_summary": dict(zip(names, map(float, fit.summaries["in"][0]))),
             "rho_out_summary": dict(zip(names, map(float, fit.summaries["out"][0]))) if has_out else nan4}
    rk_in = float(fit.rk_in[0])
    rk_out = float(fit.rk_out[0, 0]) if has_out else float("inf")
    return fit.centroids[0].astype(np.float32), float(min(rk_in, rk_out)), rk_in, rk_out, extra


# ----------------------------------------------------------------------------------------------------------
# 09_evaluate_wav_detection.py / 10_benchmark_folder_detection.py
# ----------------------------------------------------------------------------------------------------------
def load_json(path: Path) -> Dict[str, Any]:
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(f"No existe: {path}")
    with open(path, "r", encoding="utf-8") as f:
        return json.load(f)


def get_detector_from_config(cfg: Dict[str, Any]) -> Tuple[Dict[str, np.ndarray], Dict[str, float], float]:
    """09:113-149."""
    rd = cfg.get("radial_detector", None)
    if not isinstance(rd, dict):
        raise ValueError("config.json no contiene radial_detector (dict). Ejecuta antes 08_fit_radial_detector.py")
    cent, thr = rd.get("centroids", None), rd.get("thresholds", None)
    if not isinstance(cent, dict) or not isinstance(thr, dict):
        raise ValueError("radial_detector debe contener 'centroids' y 'thresholds' como dicts.")
    centroids = {sp: np.array(vec, dtype=np.float32) for sp, vec in cent.items()
                 if isinstance(sp, str) and isinstance(vec, list) and len(vec) > 0}
    thresholds = {sp: float(v) for sp, v in thr.items() if isinstance(sp, str)}
    if not centroids or not thresholds:
        raise ValueError("centroids/thresholds vacíos o mal formateados en config.json.")
    try:
        chunk_seconds = float(cfg.get("chunk_seconds", 5.0))
    except Exception:
        chunk_seconds = 5.0
    return centroids, thresholds, chunk_seconds


def l2(a: np.ndarray) -> float:
    """09:354-355."""
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(1, -1)
    return float(l2_norm_rows(a)[0])


def _decide_many(Z: np.ndarray, centroids: Dict[str, np.ndarray], thresholds: Dict[str, float]):
    """D2 for rows of ``Z``: ``-> [(detected, species | None, best_d)]`` (09:416-436, 10:175-199)."""
    D = Z.shape[1]
    species = [sp for sp, mu in centroids.items() if sp in thresholds and mu.shape[0] == D]
    if not species or Z.shape[0] == 0:
        return [(False, None, float("inf"))] * Z.shape[0]
    eng = _engine(144000, 0)
    cent = torch.from_numpy(np.stack([centroids[sp] for sp in species]).astype(np.float32)).to(eng.device)
    thr = torch.tensor([thresholds[sp] for sp in species], dtype=torch.float64, device=eng.device)
    prio = torch.from_numpy(priority_ranks(species, PRIORITY_ORDER)).to(eng.device)
    radii = eng.radii(_dev_rows(Z), cent)
    pred, best = eng.decide(radii, thr, prio)
    pred, best = pred.cpu().numpy(), best.cpu().numpy()
    return [(bool(p >= 0), species[p] if p >= 0 else None, float(b)) for p, b in zip(pred, best)]


class DetectorSession:
    """10:113-199: load config + encoder once, then ``predict_one`` per file (or ``predict_many``)."""

    def __init__(self, module=None, project_root: Path = Path("."), config_path: Path = Path("config.json"),
                 encoder_pt: Path = Path("model.pt"), encoder_yaml: Path = Path("model.yaml"), device: str = "cpu",
                 sr: int = 48000, n_mels: int = 64, target_frames: int = 192, fmin: float = 150.0,
                 fmax: float = 15000.0, hop_length: int = 384, n_fft: int = 2048):
        self.m = module
        self.project_root, self.config_path = Path(project_root), Path(config_path)
        self.encoder_pt, self.encoder_yaml, self.device = Path(encoder_pt), Path(encoder_yaml), device
        self.sr, self.n_mels, self.target_frames = sr, n_mels, target_frames
        self.fmin, self.fmax, self.hop_length, self.n_fft = fmin, fmax, hop_length, n_fft
        self.centroids: Dict[str, np.ndarray] = {}
        self.thresholds: Dict[str, float] = {}
        self.duration = 5.0
        self.encoder: Optional[torch.nn.Module] = None

    def load(self) -> None:
        cfg = load_json(self.config_path)
        self.centroids, self.thresholds, self.duration = get_detector_from_config(cfg)
        self.encoder = load_encoder(self.encoder_pt, self.encoder_yaml, self.project_root, self.device)

    def _mel_kw(self):
        return dict(sr=self.sr, duration=self.duration, n_mels=self.n_mels, fmin=self.fmin, fmax=self.fmax,
                    hop_length=self.hop_length, n_fft=self.n_fft, target_frames=self.target_frames)

    def predict_one(self, wav_path: Path) -> Tuple[bool, Optional[str], float]:
        if self.encoder is None:
            raise RuntimeError("Session no cargada. Llama load().")
        z = encode_wav_to_latent(self.encoder, wav_path, self.device, **self._mel_kw())
        return _decide_many(z[None], self.centroids, self.thresholds)[0]

    def predict_many(self, wav_paths: Sequence[Path]) -> List[Tuple[bool, Optional[str], float]]:
        """Batched ``predict_one``; an unreadable file yields ``(False, "ERROR", nan)`` (10:409-418)."""
        if self.encoder is None:
            raise RuntimeError("Session no cargada. Llama load().")
        Z, failed = encode_wavs_to_latents(self.encoder, wav_paths, self.device, return_failed=True, **self._mel_kw())
        good = iter(_decide_many(Z, self.centroids, self.thresholds))
        bad = set(map(str, failed))
        return [(False, "ERROR", float("nan")) if str(p) in bad else next(good) for p in wav_paths]


def detect_species(wav_path, *, config_path=None, encoder_pt=None, encoder_yaml=None, device: str = "cpu",
                   sr: int = 48000, n_mels: int = 64, target_frames: int = 192, fmin: float = 150.0,
                   fmax: float = 15000.0, hop_length: int = 384, n_fft: int = 2048,
                   encoder: Optional[torch.nn.Module] = None) -> Tuple[bool, Optional[str]]:
    """09:358-436 -> ``(True, species)`` / ``(False, None)``.  ``config_path`` / ``encoder_pt`` / ``encoder_yaml``
    default to ``<cwd>/config.json``, ``downloaded_models/model.pt`` and ``downloaded_models/model.yaml``.
    ``encoder=`` (addition) passes an already loaded module, avoiding the checkpoint reload per call (09:400)."""
    wav_p = Path(wav_path).expanduser()
    if not wav_p.is_absolute():
        wav_p = (Path.cwd() / wav_p).resolve()
    if not wav_p.exists():
        raise FileNotFoundError(f"No existe WAV: {wav_p}")
    root = Path.cwd()
    cfg = load_json(Path(config_path).expanduser().resolve() if config_path else root / "config.json")
    centroids, thresholds, duration = get_detector_from_config(cfg)
    if encoder is None:
        encoder = load_encoder(Path(encoder_pt) if encoder_pt else root / "downloaded_models" / "model.pt",
                               Path(encoder_yaml) if encoder_yaml else root / "downloaded_models" / "model.yaml",
                               root, device)
    z = encode_wav_to_latent(encoder, wav_p, device, sr=sr, duration=duration, n_mels=n_mels, fmin=fmin, fmax=fmax,
                             hop_length=hop_length, n_fft=n_fft, target_frames=target_frames)
    det, sp, _ = _decide_many(z[None], centroids, thresholds)[0]
    return det, sp


# ----------------------------------------------------------------------------------------------------------
# Row N1 -- Gaussian-MAP detector: map_detector_core.py:306-420, 08b_fit_map_detector.py:60-81,
# 09n_evaluate_wav_detection.py:51-140, 10b_benchmark_folder_detection_map.py:88-169
# ----------------------------------------------------------------------------------------------------------
from .map_fit import MapFit, inv_and_logdet, regularise_cov  # noqa: E402  (inv_and_logdet: core:306-316, host D x D)


def estimate_cov(Z: np.ndarray, eps: float, shrink: float, cov_structure: str) -> np.ndarray:
    """08b:60-81: covariance of the rows of ``Z`` (second moments on the GPU in float64) + diag / shrink / eps I."""
    Z = np.ascontiguousarray(Z, dtype=np.float32)
    n, d = Z.shape
    if n < 2:
        cov = np.eye(d, dtype=np.float32)
    else:
        eng = _engine(144000, 0)
        Zd = _dev_rows(Z)
        lab = torch.zeros(n, dtype=torch.int32, device=eng.device)
        zero = torch.zeros(1, d, dtype=torch.float32, device=eng.device)
        S = eng.cov_accumulate(Zd, lab, zero, 0).cpu().numpy()
        s1, _ = eng.centroid_accumulate(Zd, lab, 1)
        m = s1[0].cpu().numpy() / n
        cov = ((S - n * np.outer(m, m)) / (n - 1)).astype(np.float32)
    return regularise_cov(cov, float(eps), float(shrink), cov_structure)


def gaussian_logpdf_from_precision(z: np.ndarray, mu: np.ndarray, prec: np.ndarray, logdet_cov: float) -> float:
    """core:319-323 (one latent; batches go through ``Engine.map_score``)."""
    d = int(z.shape[0])
    fit = MapFit(["_"], np.zeros(1, np.int32), np.asarray(mu, np.float32)[None], np.zeros((1, d, d), np.float32),
                 np.asarray(prec, np.float32)[None], np.array([float(logdet_cov)]), np.array([1.0 - 1e-12]), None,
                 np.ones(1, np.int64))
    eng = _engine(144000, 0)
    _, best, _ = eng.map_score(_dev_rows(np.asarray(z, np.float32)[None]), fit)
    return float(best[0].item())


def get_priors_from_map_meta(cfg: Dict[str, Any], species: List[str]) -> Dict[str, float]:
    """core:326-355."""
    priors: Dict[str, float] = {}
    md = cfg.get("map_detector", {})
    meta = md.get("meta_fit", {}) if isinstance(md, dict) else {}
    per = meta.get("per_species", {}) if isinstance(meta, dict) else {}
    ok = True
    for sp in species:
        try:
            priors[sp] = float(per.get(sp, {}).get("prior"))
        except Exception:
            ok = False
            break
    if ok and priors:
        s = sum(max(0.0, v) for v in priors.values())
        if s > 0:
            priors = {k: max(0.0, v) / s for k, v in priors.items()}
        return priors
    K = len(species)
    return {sp: 1.0 / K for sp in species} if K else {}


def get_chunk_seconds_for_map(cfg: Dict[str, Any]) -> float:
    """core:358-370."""
    md = cfg.get("map_detector", {})
    if isinstance(md, dict):
        meta = md.get("meta_fit", {})
        if isinstance(meta, dict) and "chunk_seconds" in meta:
            try:
                return float(meta["chunk_seconds"])
            except Exception:
                pass
    try:
        return float(cfg.get("chunk_seconds", 5.0))
    except Exception:
        return 5.0


def read_map_detector_params(cfg: Dict[str, Any]):
    """core:373-420 -> ``(means, precisions, logdets, tau)``."""
    md = cfg.get("map_detector", None)
    if not isinstance(md, dict):
        raise ValueError("config.json no contiene map_detector (dict). Ejecuta antes 08b_fit_map_detector.py")
    if md.get("model", "") != "gaussian_map":
        raise ValueError(f"map_detector.model inesperado: {md.get('model')}")
    means_raw, prec_raw, logdet_raw = md.get("means"), md.get("precision"), md.get("logdet_cov")
    if not isinstance(means_raw, dict) or not isinstance(prec_raw, dict) or not isinstance(logdet_raw, dict):
        raise ValueError("map_detector debe contener 'means', 'precision' y 'logdet_cov' como dicts.")
    means = {sp: np.array(v, dtype=np.float32) for sp, v in means_raw.items()
             if isinstance(sp, str) and isinstance(v, list) and len(v) > 0}
    precisions: Dict[str, np.ndarray] = {}
    for sp, mat in prec_raw.items():
        if isinstance(sp, str) and isinstance(mat, list) and len(mat) > 0:
            Pm = np.array(mat, dtype=np.float32)
            if Pm.ndim != 2 or Pm.shape[0] != Pm.shape[1]:
                raise ValueError(f"precision[{sp}] debe ser matriz cuadrada, obtuve shape={Pm.shape}")
            precisions[sp] = Pm
    logdets = {sp: float(v) for sp, v in logdet_raw.items() if isinstance(sp, str)}
    tau = md.get("tau", None)
    if not means or not precisions or not logdets:
        raise ValueError("means/precision/logdet_cov vacíos o mal formateados en config.json.")
    return means, precisions, logdets, (float(tau) if tau is not None else None)


def _map_fit_from_params(means, precisions, logdets, priors, tau, D: int) -> Optional[MapFit]:
    species = sorted(set(means) & set(precisions) & set(logdets))
    species = [sp for sp in species if means[sp].shape[0] == D and precisions[sp].shape == (D, D)]   # 09n:120-123
    if not species:
        return None
    pri = np.array([float(priors.get(sp, 1e-12)) for sp in species], dtype=np.float64)
    return MapFit(species, np.arange(len(species), dtype=np.int32), np.stack([means[sp] for sp in species]),
                  np.zeros((len(species), D, D), np.float32), np.stack([precisions[sp] for sp in species]),
                  np.array([logdets[sp] for sp in species], dtype=np.float64), pri, tau,
                  np.zeros(len(species), np.int64))


def _decide_map_many(Z: np.ndarray, means, precisions, logdets, priors, tau):
    fit = _map_fit_from_params(means, precisions, logdets, priors, tau, Z.shape[1])
    if fit is None or Z.shape[0] == 0:
        return [(False, None, -float("inf"))] * Z.shape[0]
    eng = _engine(144000, 0)
    pred, best, _ = eng.map_score(_dev_rows(Z), fit)
    pred, best = pred.cpu().numpy(), best.cpu().numpy()
    return [(bool(p >= 0), fit.species[p] if p >= 0 else None, float(b)) for p, b in zip(pred, best)]


class MapDetectorSession:
    """10b:88-169: MAP analogue of :class:`DetectorSession`."""

    def __init__(self, project_root: Path = Path("."), config_path: Path = Path("config.json"),
                 encoder_pt: Path = Path("model.pt"), encoder_yaml: Path = Path("model.yaml"), device: str = "cpu",
                 sr: int = 48000, n_mels: int = 64, target_frames: int = 192, fmin: float = 150.0,
                 fmax: float = 15000.0, hop_length: int = 384, n_fft: int = 2048):
        self.project_root, self.config_path = Path(project_root), Path(config_path)
        self.encoder_pt, self.encoder_yaml, self.device = Path(encoder_pt), Path(encoder_yaml), device
        self.sr, self.n_mels, self.target_frames = sr, n_mels, target_frames
        self.fmin, self.fmax, self.hop_length, self.n_fft = fmin, fmax, hop_length, n_fft
        self.means, self.precisions, self.logdets, self.priors = {}, {}, {}, {}
        self.tau: Optional[float] = None
        self.species: List[str] = []
        self.duration = 5.0
        self.encoder: Optional[torch.nn.Module] = None

    def load(self) -> None:
        cfg = load_json(self.config_path)
        self.set_params(cfg)
        self.encoder = load_encoder(self.encoder_pt, self.encoder_yaml, self.project_root, self.device)

    def set_params(self, cfg: Dict[str, Any]) -> None:
        self.means, self.precisions, self.logdets, self.tau = read_map_detector_params(cfg)
        self.species = sorted(set(self.means) & set(self.precisions) & set(self.logdets))
        if not self.species:
            raise RuntimeError("map_detector inconsistente: no hay intersección entre means/precision/logdet_cov.")
        self.priors = get_priors_from_map_meta(cfg, self.species)
        self.duration = float(get_chunk_seconds_for_map(cfg))

    def _mel_kw(self):
        return dict(sr=self.sr, duration=self.duration, n_mels=self.n_mels, fmin=self.fmin, fmax=self.fmax,
                    hop_length=self.hop_length, n_fft=self.n_fft, target_frames=self.target_frames)

    def predict_one(self, wav_path: Path) -> Tuple[bool, Optional[str], float]:
        z = encode_wav_to_latent(self.encoder, wav_path, self.device, **self._mel_kw())
        return _decide_map_many(z[None], self.means, self.precisions, self.logdets, self.priors, self.tau)[0]

    def predict_many(self, wav_paths: Sequence[Path]) -> List[Tuple[bool, Optional[str], float]]:
        Z, failed = encode_wavs_to_latents(self.encoder, wav_paths, self.device, return_failed=True, **self._mel_kw())
        good = iter(_decide_map_many(Z, self.means, self.precisions, self.logdets, self.priors, self.tau))
        bad = set(map(str, failed))
        return [(False, "ERROR", float("nan")) if str(p) in bad else next(good) for p in wav_paths]


def detect_species_map(wav_path, *, config_path=None, encoder_pt=None, encoder_yaml=None, device: str = "cpu",
                       sr: int = 48000, n_mels: int = 64, target_frames: int = 192, fmin: float = 150.0,
                       fmax: float = 15000.0, hop_length: int = 384, n_fft: int = 2048,
                       encoder: Optional[torch.nn.Module] = None) -> Tuple[bool, Optional[str], float]:
    """09n:51-140 -> ``(detected, species | None, best_score)``."""
    wav_p = Path(wav_path).expanduser()
    if not wav_p.is_absolute():
        wav_p = (Path.cwd() / wav_p).resolve()
    if not wav_p.exists():
        raise FileNotFoundError(f"No existe WAV: {wav_p}")
    root = Path.cwd()
    sess = MapDetectorSession(root, Path(config_path).expanduser().resolve() if config_path else root / "config.json",
                              Path(encoder_pt) if encoder_pt else root / "models" / "model.pt",
                              Path(encoder_yaml) if encoder_yaml else root / "models" / "model.yaml", device, sr, n_mels,
                              target_frames, fmin, fmax, hop_length, n_fft)
    sess.set_params(load_json(sess.config_path))
    sess.encoder = encoder if encoder is not None else load_encoder(sess.encoder_pt, sess.encoder_yaml, root, device)
    return sess.predict_one(wav_p)


# ----------------------------------------------------------------------------------------------------------
# Row N4b -- 07_encode_wav_to_latent.py: output unpacking, guarded forward and the target_frames probe for foreign
# encoders (07:195-199, :264-409).  Host logic around an arbitrary nn.Module; only wav_to_mel runs on the GPU library.
# ----------------------------------------------------------------------------------------------------------
def find_first_linear(module: torch.nn.Module) -> torch.nn.Linear:
    """07:195-199."""
    for m in module.modules():
        if isinstance(m, torch.nn.Linear):
            return m
    raise RuntimeError("No encontré ningún nn.Linear dentro del encoder.")


_VECTOR_KEYS_07 = ("z", "latent", "mu", "mean", "embedding", "enc")        # 07:278 (core:283 has no "enc")


def extract_vector(out: Any) -> torch.Tensor:
    """07:264-297: encoder output -> ``[B, D]`` tensor.  Tensor as is; list/tuple -> first tensor (a VAE's ``mu``); dict ->
    first present of ``z, latent, mu, mean, embedding, enc`` else the first tensor value; ``[B, T', C]`` -> mean over
    ``T'`` (an empty ``T'`` gives ``[B, 0]``); more than two dims -> flattened per row."""
    t: Optional[torch.Tensor]
    if isinstance(out, torch.Tensor):
        t = out
    elif isinstance(out, (list, tuple)):
        t = next((o for o in out if isinstance(o, torch.Tensor)), None)
        if t is None:
            raise ValueError("Salida list/tuple sin tensors.")
    elif isinstance(out, dict):
        t = next((out[k] for k in _VECTOR_KEYS_07 if isinstance(out.get(k), torch.Tensor)), None)
        if t is None:
            t = next((v for v in out.values() if isinstance(v, torch.Tensor)), None)
        if t is None:
            raise ValueError("Salida dict sin tensors.")
    else:
        raise ValueError(f"No sé interpretar salida: {type(out)}")
    if t.ndim == 3:
        if t.shape[1] == 0:
            return t[:, :0]
        t = t.mean(dim=1)
    if t.ndim > 2:
        t = t.view(t.shape[0], -1)
    return t


def try_forward(encoder: torch.nn.Module, x: torch.Tensor) -> Tuple[bool, Optional[torch.Tensor], Optional[str]]:
    """07:300-310 -> ``(ok, vec, err)``; any exception of the forward or the unpacking becomes ``(False, None, str(e))``."""
    try:
        with torch.no_grad():
            vec = extract_vector(encoder(x))
    except Exception as e:      # noqa: BLE001 -- the reference reports, it does not raise
        return False, None, str(e)
    if vec.numel() == 0 or vec.shape[1] == 0:
        return False, None, "Salida vacía (numel=0)."
    return True, vec, None


@torch.no_grad()
def probe_linear_input_shape(encoder: torch.nn.Module, linear: torch.nn.Linear, x: torch.Tensor):
    """07:317-352 -> ``(ok, (N, F) | None, err | None)``: the shape entering ``linear`` during ``encoder(x)``, with every
    leading dimension collapsed into N.  A forward that fails *after* the hook fired still reports the shape (with the
    error text)."""
    seen: Dict[str, Any] = {"shape": None}

    def pre_hook(_mod, inputs):
        inp = inputs[0] if inputs else None
        if isinstance(inp, torch.Tensor) and inp.ndim >= 2 and inp.shape[-1] > 0:
            feat = int(inp.shape[-1])
            seen["shape"] = (int(inp.numel() // feat), feat)

    handle = linear.register_forward_pre_hook(pre_hook)
    err: Optional[str] = None
    try:
        encoder(x)
    except Exception as e:      # noqa: BLE001
        err = str(e)
    finally:
        handle.remove()
    if seen["shape"] is None:
        return False, None, err if err is not None else "No pude capturar la entrada al Linear (hook no disparó)."
    return True, seen["shape"], err


def auto_find_frames_with_hook(encoder: torch.nn.Module, wav_path: Path, device, sr: int, duration: float, n_mels: int,
                               fmin: float, fmax: float, hop_length: int, n_fft: int, start_frames: int, max_frames: int,
                               step: int) -> int:
    """07:355-409: the first ``target_frames`` in ``start, start + step, ... <= max_frames`` (start >= 8) for which the
    tensor reaching the encoder's first ``nn.Linear`` has the width that layer expects and is not empty.  The log-mel
    features are computed once at the largest size the loop can ask for and cropped / padded per candidate exactly as
    ``crop_or_pad_time`` does, instead of one feature pass per candidate."""
    linear = find_first_linear(encoder)
    want = int(linear.in_features)
    start = max(8, int(start_frames))
    step = max(1, int(step))
    max_frames = max(start, int(max_frames))
    dev = torch.device(device) if not isinstance(device, torch.device) else device
    chunk_len = int(sr * duration)
    n_native = 1 + chunk_len // hop_length
    # one GPU pass: ask for every frame (target_frames = native count), crop / pad on the host per candidate
    full = wav_to_mel(Path(wav_path), sr=sr, duration=duration, n_mels=n_mels, fmin=fmin, fmax=fmax, hop_length=hop_length,
                      n_fft=n_fft, target_frames=n_native).numpy()
    for frames in range(start, max_frames + 1, step):
        mel = torch.from_numpy(crop_or_pad_time(full, frames).astype(np.float32))
        x = mel.T.unsqueeze(0).unsqueeze(0).to(dev)                  # [1, 1, T, M] (07:395)
        ok, shp, _ = probe_linear_input_shape(encoder, linear, x)
        if ok and shp is not None and shp[1] == want and shp[0] > 0:
            return frames
    raise SystemExit("❌ No encontré un target_frames válido usando el hook.\n"
                     f"Probé frames desde {start} hasta {max_frames} step={step}.\n"
                     "Tip: sube --auto-max-frames (ej. 4096) o baja --auto-step (ej. 1).")


def encode_wav_report(wav_path: Path, encoder: torch.nn.Module, *, device="cuda", sr: int = 48000, duration: float = 3.0,
                      n_mels: int = 64, fmin: float = 150.0, fmax: float = 15000.0, hop_length: int = 384,
                      n_fft: int = 2048, target_frames: int = 192, auto_frames: bool = False, auto_max_frames: int = 512,
                      auto_step: int = 8, jsonl: bool = False, precision: int = 6, log=print) -> np.ndarray:
    """07 ``main`` after the encoder is loaded (07:472-527): optional frame search, features on the GPU library, the
    encoder module's own forward (any architecture), and the script's two output formats."""
    import json as _json
    wav_path = Path(wav_path)
    frames = auto_find_frames_with_hook(encoder, wav_path, device, sr, duration, n_mels, fmin, fmax, hop_length, n_fft,
                                        target_frames, auto_max_frames, auto_step) if auto_frames else target_frames
    log(f"✅ target_frames usado: {frames}")
    mel = wav_to_mel(wav_path, sr=sr, duration=duration, n_mels=n_mels, fmin=fmin, fmax=fmax, hop_length=hop_length,
                     n_fft=n_fft, target_frames=frames)
    dev = torch.device(device) if not isinstance(device, torch.device) else device
    x = mel.T.unsqueeze(0).unsqueeze(0).to(dev)
    ok, vec, err = try_forward(encoder.to(dev), x)
    if not ok or vec is None:
        raise SystemExit(f"❌ Forward falló incluso con target_frames={frames}. Error: {err}")
    v = vec.detach().cpu().numpy()[0]
    if jsonl:
        log(_json.dumps({"wav": str(wav_path), "latent_dim": int(v.size), "vector": v.astype(float).tolist()},
                        ensure_ascii=False))
    else:
        with np.printoptions(precision=int(precision), suppress=True, linewidth=180):
            log(f"✅ Latent dim: {v.size}")
            log(str(v))
    return v
