"""Seeded synthetic chunks (the reference ships no audio; definition in SURVEY.md section 8d).

Mono float32 at 48 kHz; species ``k = i % K`` round robin; chunk ``i`` = low-passed noise
(sigma log-uniform 1e-3..0.1) + species-specific AM tone bursts (carrier 1.1 / 1.8 / 2.6 / 3.4 kHz
+-5 %, pulse rate 4..12 Hz, amplitude log-uniform 0.01..0.5); ~1 % silent chunks (sigma = 1e-5,
exercises the ``rms < rms_min`` gate of 00_normalize_dataset_rms.py:32-34) and ~1 % hot chunks
(amplitude 2.0, exercises the clip at 00_normalize_dataset_rms.py:37).

Pure torch, vectorised over the batch, runs on CPU (tests, oracle) or on the GPU (bench input
generation -- setup only, never inside a timed region).
"""
from __future__ import annotations

import math
from typing import Tuple

import torch

CARRIERS_HZ = (1100.0, 1800.0, 2600.0, 3400.0)


def make_chunks(n: int, length: int = 144000, *, sr: int = 48000, n_species: int = 4, seed: int = 123,
                first_index: int = 0, device: str | torch.device = "cpu",
                special_every: int = 100) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns ``(x [n, length] float32, label [n] int32)``.  Deterministic in ``(seed, first_index)``
    per device type; chunk ``i`` has label ``(first_index + i) % n_species``."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed + 7919 * first_index)
    idx = torch.arange(first_index, first_index + n, device=dev)
    label = (idx % n_species).to(torch.int32)

    def u(lo: float, hi: float) -> torch.Tensor:
        return lo + (hi - lo) * torch.rand(n, 1, generator=g, device=dev)

    sigma = torch.exp(u(math.log(1e-3), math.log(0.1)))
    amp = torch.exp(u(math.log(0.01), math.log(0.5)))
    carrier = torch.tensor(CARRIERS_HZ, device=dev)[(label % len(CARRIERS_HZ)).long()].view(n, 1) * u(0.95, 1.05)
    rate = u(4.0, 12.0)
    phase = u(0.0, 2 * math.pi)

    noise = torch.randn(n, length + 3, generator=g, device=dev)
    noise = 0.5 * (noise[:, 3:] + noise[:, 2:-1] + noise[:, 1:-2] + noise[:, :-3])   # 4-tap low-pass
    t = torch.arange(length, device=dev, dtype=torch.float32).view(1, length) / float(sr)
    env = torch.clamp(torch.sin(2 * math.pi * rate * t + phase), min=0.0) ** 2
    x = sigma * noise + amp * env * torch.sin(2 * math.pi * carrier * t)

    if special_every > 0:
        silent = (idx % special_every) == (special_every - 3)
        hot = (idx % special_every) == (special_every - 7)
        if silent.any():
            x[silent] = 1e-5 * noise[silent]
        if hot.any():
            x[hot] = x[hot] * (2.0 / x[hot].abs().amax(dim=1, keepdim=True).clamp_min(1e-9))
    return x.to(torch.float32).contiguous(), label


def write_wav_tree(root, species, n_per_class: int, length: int = 144000, *, sr: int = 48000, seed: int = 123,
                   special_every: int = 0):
    """``<root>/<species>/<species>_<iii>.wav`` (PCM_16 mono): the folder layout 00 / 08 / 10 walk
    (00_normalize_dataset_rms.py:41-57, 08_fit_radial_detector.py:461-486, 10_benchmark_folder_detection.py:387-395).
    Always generated on the CPU so that the files are identical wherever they are made."""
    import wave
    from pathlib import Path

    import numpy as np

    root = Path(root)
    K = len(species)
    x, label = make_chunks(n_per_class * K, length, sr=sr, n_species=K, seed=seed, special_every=special_every)
    pcm = torch.clamp(torch.round(x * 32767.0), -32768, 32767).to(torch.int16).numpy()
    label = label.numpy()
    files = []
    for k, sp in enumerate(species):
        d = root / sp
        d.mkdir(parents=True, exist_ok=True)
        for j, i in enumerate(np.nonzero(label == k)[0]):
            path = d / f"{sp}_{j:03d}.wav"
            with wave.open(str(path), "wb") as w:
                w.setnchannels(1)
                w.setsampwidth(2)
                w.setframerate(int(sr))
                w.writeframes(pcm[i].astype("<i2").tobytes())
            files.append(path)
    return files
