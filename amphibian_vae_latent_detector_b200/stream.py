"""Long-recording front-end (BASELINE.json configs[4]).

The reference has no sliding-window code: every consumer truncates / zero-pads a file to ``chunk_seconds``
(map_detector_core.py:212-217) and the chunking of field recordings happens outside the repository (README.md:26-29).
The semantics defined for a long WAV are therefore the ones that reproduce that workflow: the
recording is cut into windows of ``chunk_seconds`` (hop = window unless ``hop_seconds`` is given), the last partial
window is zero-padded on the right as ``wav_to_mel`` pads a short file, and every window goes through exactly what one
chunk file goes through -- ``rms_normalize`` + the PCM_16 write/read of ``process_folder`` (00:29-57), log-mel, encoder
mean, radial decision (09:416-436) or, 09n / 10b style, Gaussian-MAP decision (``detect_long_wav_map``) -- so window ``i``
of the stream gives the same answer as the chunk file holding the same samples.

Data path: PCM_16 samples are sliced straight out of the (memory-mapped) WAV payload into pinned host slabs by a
filler thread while the previous slab is inside ``avld_encode_detect_host_pcm16`` (which itself double-buffers
``max_batch``-window pieces between its copy and compute streams), so disk read, H2D and kernels overlap.
"""
from __future__ import annotations

import os
import threading
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np
import torch

from . import reference_api as api
from .engine import priority_ranks


@dataclass
class WindowResult:
    start_s: float
    detected: bool
    species: Optional[str]
    best_distance: float
    normalised: bool          # False = the window was below the rms_min gate (00:32-34) and passed through unscaled


@dataclass
class MapWindowResult:
    """One window decided by the Gaussian-MAP detector (09n:114-140): ``best_score`` is the largest class score, reported
    for rejected windows too."""
    start_s: float
    detected: bool
    species: Optional[str]
    best_score: float
    normalised: bool


def window_starts(n_samples: int, window_len: int, hop_len: int) -> np.ndarray:
    """Start sample of every window: 0, hop, 2 hop, ... while start < n_samples (an empty recording has none)."""
    if n_samples <= 0:
        return np.zeros(0, dtype=np.int64)
    return np.arange(0, n_samples, hop_len, dtype=np.int64)


def open_pcm16_mono(path, sr: int) -> np.ndarray:
    """Mono PCM_16 WAV -> int16 sample array, memory-mapped (nothing is read until it is sliced).  Raises ``ValueError``
    for any other sample format / channel count (the caller then decodes as ``librosa.load`` would).

    A day of 48 kHz PCM_16 audio is 8.3 GB, beyond the 32-bit sizes of RIFF, so the 64-bit forms recorders and libsndfile
    write are accepted too: ``RF64`` / ``BW64`` files (true data size in the ``ds64`` chunk, the 32-bit field holds
    0xFFFFFFFF) and plain ``RIFF`` files whose data size field is the 0xFFFFFFFF / 0 placeholder of a streamed write --
    both mean "to the end of the file"."""
    import struct
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] not in (b"RIFF", b"RF64", b"BW64") or head[8:12] != b"WAVE":
            raise RuntimeError(f"{path}: not a RIFF/WAVE file")
        fmt = None
        data_size64 = None
        while True:
            hdr = f.read(8)
            if len(hdr) < 8:
                raise RuntimeError(f"{path}: no data chunk")
            name, size = hdr[:4], struct.unpack("<I", hdr[4:])[0]
            if name == b"fmt ":
                fmt = struct.unpack("<HHIIHH", f.read(16))
                f.seek(size - 16 + (size & 1), 1)
            elif name == b"ds64":
                body = f.read(size + (size & 1))
                if len(body) >= 16:
                    data_size64 = struct.unpack("<Q", body[8:16])[0]                 # riffSize, dataSize, sampleCount, ...
            elif name == b"data":
                offset = f.tell()
                break
            else:
                f.seek(size + (size & 1), 1)
    if fmt is None:
        raise RuntimeError(f"{path}: data chunk before fmt chunk")
    tag, nch, rate, _, _, bits = fmt
    if rate != sr:
        raise RuntimeError(f"{path}: sample rate {rate} != {sr}; resampling is not implemented on this path")
    if tag != 1 or nch != 1 or bits != 16:
        raise ValueError("not mono PCM_16")
    remaining = Path(path).stat().st_size - offset
    if size == 0xFFFFFFFF or (size == 0 and remaining > 0):
        size = data_size64 if data_size64 else remaining
    n = min(size, remaining) // 2
    if n == 0:
        return np.zeros(0, dtype="<i2")
    return np.memmap(str(path), dtype="<i2", mode="r", offset=offset, shape=(n,))


def _fill_rows(dst: np.ndarray, pcm: np.ndarray, starts: np.ndarray, window_len: int) -> None:
    n = pcm.shape[0]
    if starts.shape[0] == 0:
        return
    hop_is_window = starts.shape[0] > 1 and int(starts[1] - starts[0]) == window_len
    last_full = int(np.searchsorted(starts + window_len, n, side="right"))       # windows that lie fully inside
    if hop_is_window and last_full > 0:
        dst[:last_full] = pcm[starts[0]:starts[0] + last_full * window_len].reshape(last_full, window_len)
    else:
        for i in range(last_full):
            dst[i] = pcm[starts[i]:starts[i] + window_len]
    for i in range(last_full, starts.shape[0]):                                   # right zero-pad (core:214-215)
        m = max(0, n - int(starts[i]))
        dst[i, :m] = pcm[starts[i]:starts[i] + m]
        dst[i, m:] = 0


_POOL = ThreadPoolExecutor(max_workers=max(2, min(8, (os.cpu_count() or 4) // 2)))
_PINNED: dict = {}          # (rows, window_len) -> two pinned int16 buffers, reused across calls (page-locking is slow)


def _fill_slab(dst: np.ndarray, pcm: np.ndarray, starts: np.ndarray, window_len: int) -> None:
    """numpy's copy releases the GIL: a slab is filled by several threads, each a contiguous block of rows
    (one memcpy stream per thread; a single thread moves ~3.5 GB/s, far below what the H2D engine takes)."""
    rows = starts.shape[0]
    parts = _POOL._max_workers
    if rows < 4 * parts:
        _fill_rows(dst, pcm, starts, window_len)
        return
    step = (rows + parts - 1) // parts
    futs = [_POOL.submit(_fill_rows, dst[i:i + step], pcm, starts[i:i + step], window_len) for i in range(0, rows, step)]
    for f in futs:
        f.result()


def iter_slabs(pcm: np.ndarray, window_len: int, hop_len: int, slab_windows: int) -> Iterator[Tuple[np.ndarray, torch.Tensor]]:
    """Yields ``(starts, pinned int16 [m, window_len])``; the next slab is filled by threads while the caller works."""
    starts = window_starts(pcm.shape[0], window_len, hop_len)
    if starts.shape[0] == 0:
        return
    pin = torch.cuda.is_available()
    rows = min(slab_windows, starts.shape[0])
    key = (rows, window_len, pin)
    bufs = _PINNED.get(key)
    if bufs is None:
        if len(_PINNED) >= 2:       # two slab geometries at most (a 2048-window slab of 5 s PCM_16 windows is ~1 GB, twice)
            _PINNED.clear()
        bufs = [torch.empty(rows, window_len, dtype=torch.int16, pin_memory=pin) for _ in range(2)]
        _PINNED[key] = bufs
    pieces = [starts[i:i + slab_windows] for i in range(0, starts.shape[0], slab_windows)]

    def fill(k):
        _fill_slab(bufs[k & 1].numpy()[:pieces[k].shape[0]], pcm, pieces[k], window_len)

    fill(0)
    for k, piece in enumerate(pieces):
        t = None
        if k + 1 < len(pieces):
            t = threading.Thread(target=fill, args=(k + 1,))
            t.start()
        yield piece, bufs[k & 1][:piece.shape[0]]
        if t is not None:
            t.join()


def detect_pcm16_stream(pcm: np.ndarray, encoder: torch.nn.Module, centroids: Dict[str, np.ndarray],
                        thresholds: Dict[str, float], *, sr: int = 48000, window_seconds: float = 5.0,
                        hop_seconds: Optional[float] = None, slab_windows: int = 2048, device=0,
                        n_mels: int = 64, fmin: float = 150.0, fmax: float = 15000.0, hop_length: int = 384,
                        n_fft: int = 2048, target_frames: int = 192) -> List[WindowResult]:
    """``pcm`` int16 [n_samples] (array or memmap) -> one :class:`WindowResult` per window."""
    window_len = int(sr * window_seconds)
    hop_len = window_len if hop_seconds is None else int(sr * hop_seconds)
    if hop_len <= 0:
        raise ValueError("hop_seconds must be positive")
    mel_kw = dict(sr=sr, n_mels=n_mels, fmin=fmin, fmax=fmax, hop_length=hop_length, n_fft=n_fft, target_frames=target_frames)
    n_win = window_starts(pcm.shape[0], window_len, hop_len).shape[0]
    eng = api._engine_with_encoder(encoder, window_len, device, max_batch=64 if n_win <= 256 else 1024, **mel_kw)
    D = eng.latent_dim
    species = [sp for sp, mu in centroids.items() if sp in thresholds and mu.shape[0] == D]     # 09:419-423
    out: List[WindowResult] = []
    if not species:
        return [WindowResult(float(s) / sr, False, None, float("inf"), True)
                for s in window_starts(pcm.shape[0], window_len, hop_len)]
    cent = np.stack([centroids[sp] for sp in species]).astype(np.float32)
    thr = np.array([thresholds[sp] for sp in species], dtype=np.float64)
    prio = priority_ranks(species, api.PRIORITY_ORDER)
    for starts, slab in iter_slabs(pcm, window_len, hop_len, slab_windows):
        pred, best, ok, _ = eng.encode_detect_host(slab, cent, thr, prio, pcm16=True)
        for s, p, b, o in zip(starts, pred, best, ok):
            out.append(WindowResult(float(s) / sr, bool(p >= 0), species[p] if p >= 0 else None, float(b), bool(o)))
    return out


def detect_long_wav(wav_path, *, config_path, encoder: torch.nn.Module, window_seconds: Optional[float] = None,
                    hop_seconds: Optional[float] = None, sr: int = 48000, slab_windows: int = 4096, device=0,
                    **mel_kw) -> List[WindowResult]:
    """Windows of ``chunk_seconds`` (config.json, default 5.0: 09:144-147) over a long mono PCM_16 WAV, decided with the
    config's radial detector (``get_detector_from_config``, 09:113-149)."""
    cfg = api.load_json(Path(config_path))
    centroids, thresholds, duration = api.get_detector_from_config(cfg)
    win = float(duration if window_seconds is None else window_seconds)
    try:
        pcm = open_pcm16_mono(wav_path, sr)
    except ValueError:
        # other sample formats / channel counts: decode as librosa.load would (float32 samples); the windows then take the
        # float path of the library, which applies the same normalise + PCM_16 round trip as the chunk-file workflow
        y = api.load_wav(wav_path, sr)
        return _detect_float_stream(y, encoder, centroids, thresholds, sr=sr, window_seconds=win, hop_seconds=hop_seconds,
                                    slab_windows=slab_windows, device=device, **mel_kw)
    return detect_pcm16_stream(pcm, encoder, centroids, thresholds, sr=sr, window_seconds=win, hop_seconds=hop_seconds,
                               slab_windows=slab_windows, device=device, **mel_kw)


def _detect_float_stream(y: np.ndarray, encoder, centroids, thresholds, *, sr, window_seconds, hop_seconds, slab_windows,
                         device, **mel_kw) -> List[WindowResult]:
    window_len = int(sr * window_seconds)
    hop_len = window_len if hop_seconds is None else int(sr * hop_seconds)
    kw = dict(sr=sr, n_mels=64, fmin=150.0, fmax=15000.0, hop_length=384, n_fft=2048, target_frames=192)
    kw.update(mel_kw)
    eng = api._engine_with_encoder(encoder, window_len, device, **kw)
    species = [sp for sp, mu in centroids.items() if sp in thresholds and mu.shape[0] == eng.latent_dim]
    starts = window_starts(y.shape[0], window_len, hop_len)
    out: List[WindowResult] = []
    if not species:
        return [WindowResult(float(s) / sr, False, None, float("inf"), True) for s in starts]
    cent = np.stack([centroids[sp] for sp in species]).astype(np.float32)
    thr = np.array([thresholds[sp] for sp in species], dtype=np.float64)
    prio = priority_ranks(species, api.PRIORITY_ORDER)
    for i in range(0, starts.shape[0], slab_windows):
        piece = starts[i:i + slab_windows]
        slab = np.zeros((piece.shape[0], window_len), dtype=np.float32)
        for j, s in enumerate(piece):
            m = min(window_len, y.shape[0] - int(s))
            slab[j, :m] = y[s:s + m]
        pred, best, ok, _ = eng.encode_detect_host(slab, cent, thr, prio, pcm16=True)
        for s, p, b, o in zip(piece, pred, best, ok):
            out.append(WindowResult(float(s) / sr, bool(p >= 0), species[p] if p >= 0 else None, float(b), bool(o)))
    return out


# ----------------------------------------------------------------------------------------------------------
# The same windows decided 09n / 10b style (BASELINE.json configs[4]): Gaussian-MAP score, argmax, tau
# ----------------------------------------------------------------------------------------------------------
def _map_windows(slabs, eng, fit, sr: int) -> List[MapWindowResult]:
    """``slabs`` yields ``(starts, host slab [m, window_len])``; the encode half of the fused host call produces the
    latent means (its radial half runs against a single dummy centroid and is ignored), ``avld_map_score`` decides."""
    D = eng.latent_dim
    zero, far, prio = np.zeros((1, D), np.float32), np.array([np.inf]), np.zeros(1, np.int32)
    out: List[MapWindowResult] = []
    for starts, slab in slabs:
        _, _, ok, mu = eng.encode_detect_host(slab, zero, far, prio, pcm16=True, want_mu=True)
        if fit is None:                                                            # no usable class (09n:142-143)
            out += [MapWindowResult(float(s) / sr, False, None, -float("inf"), bool(o)) for s, o in zip(starts, ok)]
            continue
        pred, best, _ = eng.map_score(torch.from_numpy(mu).to(eng.device), fit)
        for s, p, b, o in zip(starts, pred.cpu().numpy(), best.cpu().numpy(), ok):
            out.append(MapWindowResult(float(s) / sr, bool(p >= 0), fit.species[p] if p >= 0 else None, float(b), bool(o)))
    return out


def detect_pcm16_stream_map(pcm: np.ndarray, encoder: torch.nn.Module, means: Dict[str, np.ndarray],
                            precisions: Dict[str, np.ndarray], logdets: Dict[str, float], priors: Dict[str, float],
                            tau: Optional[float], *, sr: int = 48000, window_seconds: float = 5.0,
                            hop_seconds: Optional[float] = None, slab_windows: int = 2048, device=0, n_mels: int = 64,
                            fmin: float = 150.0, fmax: float = 15000.0, hop_length: int = 384, n_fft: int = 2048,
                            target_frames: int = 192) -> List[MapWindowResult]:
    """``pcm`` int16 [n_samples] (array or memmap) -> one :class:`MapWindowResult` per window; the parameters are those
    ``read_map_detector_params`` / ``get_priors_from_map_meta`` return (core:326-420)."""
    window_len = int(sr * window_seconds)
    hop_len = window_len if hop_seconds is None else int(sr * hop_seconds)
    if hop_len <= 0:
        raise ValueError("hop_seconds must be positive")
    mel_kw = dict(sr=sr, n_mels=n_mels, fmin=fmin, fmax=fmax, hop_length=hop_length, n_fft=n_fft, target_frames=target_frames)
    n_win = window_starts(pcm.shape[0], window_len, hop_len).shape[0]
    eng = api._engine_with_encoder(encoder, window_len, device, max_batch=64 if n_win <= 256 else 1024, **mel_kw)
    fit = api._map_fit_from_params(means, precisions, logdets, priors, tau, eng.latent_dim)
    return _map_windows(iter_slabs(pcm, window_len, hop_len, slab_windows), eng, fit, sr)


def detect_long_wav_map(wav_path, *, config_path, encoder: torch.nn.Module, window_seconds: Optional[float] = None,
                        hop_seconds: Optional[float] = None, sr: int = 48000, slab_windows: int = 4096, device=0,
                        **mel_kw) -> List[MapWindowResult]:
    """Windows of ``map_detector.meta_fit.chunk_seconds`` (core:358-370) over a long WAV, decided with the config's
    Gaussian-MAP detector as ``09n_evaluate_wav_detection.py`` decides one chunk file."""
    cfg = api.load_json(Path(config_path))
    means, precisions, logdets, tau = api.read_map_detector_params(cfg)
    species = sorted(set(means) & set(precisions) & set(logdets))
    if not species:
        raise RuntimeError("map_detector inconsistente: no hay intersección entre means/precision/logdet_cov.")     # 09n:93-94
    priors = api.get_priors_from_map_meta(cfg, species)
    win = float(api.get_chunk_seconds_for_map(cfg) if window_seconds is None else window_seconds)
    try:
        pcm = open_pcm16_mono(wav_path, sr)
    except ValueError:                       # other sample formats / channel counts: decoded as librosa.load would
        y = api.load_wav(wav_path, sr)
        window_len = int(sr * win)
        hop_len = window_len if hop_seconds is None else int(sr * hop_seconds)
        kw = dict(sr=sr, n_mels=64, fmin=150.0, fmax=15000.0, hop_length=384, n_fft=2048, target_frames=192)
        kw.update(mel_kw)
        eng = api._engine_with_encoder(encoder, window_len, device, **kw)
        fit = api._map_fit_from_params(means, precisions, logdets, priors, tau, eng.latent_dim)
        starts = window_starts(y.shape[0], window_len, hop_len)

        def float_slabs():
            for i in range(0, starts.shape[0], slab_windows):
                piece = starts[i:i + slab_windows]
                slab = np.zeros((piece.shape[0], window_len), dtype=np.float32)
                for j, st in enumerate(piece):
                    m = min(window_len, y.shape[0] - int(st))
                    slab[j, :m] = y[st:st + m]
                yield piece, slab

        return _map_windows(float_slabs(), eng, fit, sr)
    return detect_pcm16_stream_map(pcm, encoder, means, precisions, logdets, priors, tau, sr=sr, window_seconds=win,
                                   hop_seconds=hop_seconds, slab_windows=slab_windows, device=device, **mel_kw)
