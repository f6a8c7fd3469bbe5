#!/usr/bin/env python
"""BASELINE.json configs[4]: 24 h of synthetic mono PCM_16 audio at 48 kHz (28 800 windows of 3 s) through
stream.detect_pcm16_stream on one GPU: pageable int16 samples -> pinned slabs (filler thread) -> H2D -> kernels ->
decisions.  Prints one JSON line; the first windows are checked against the chunk API."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from amphibian_vae_latent_detector_b200 import stream, synth  # noqa: E402
from amphibian_vae_latent_detector_b200 import reference_api as api  # noqa: E402
from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder  # noqa: E402
from amphibian_vae_latent_detector_b200.engine import priority_ranks  # noqa: E402

L, HOURS = 144000, float(sys.argv[1]) if len(sys.argv) > 1 else 24.0
n_win = int(HOURS * 3600 / 3)
species = ["Batrachyla_leptopus", "Batrachyla_taeniata", "Calyptocephalella_gayi", "Pleurodema_thaul"]
enc = build_standin_encoder(seed=123)
x, lab = synth.make_chunks(256, L, seed=123, device="cuda")
block = torch.clamp(torch.round(x * 32767.0), -32768, 32767).to(torch.int16).cpu().numpy().reshape(-1)
pcm = np.tile(block, (n_win + 255) // 256)[: n_win * L]            # pageable host memory, as a decoded recording would be
eng = api._engine_with_encoder(enc, L, 0, max_batch=1024, sr=48000, n_mels=64, fmin=150.0, fmax=15000.0, hop_length=384,
                               n_fft=2048, target_frames=192)
Z, ok = eng.encode(x, pcm16=True)
fit = eng.fit_radial(Z, lab, 4, 0.95, [0.25])
cents = {sp: np.nan_to_num(fit.centroids[i]) for i, sp in enumerate(species)}
thr = {sp: float(fit.rk[0, i]) for i, sp in enumerate(species)}
stream.detect_pcm16_stream(pcm[: 2048 * L], enc, cents, thr, window_seconds=3.0)          # warm-up
t0 = time.perf_counter()
res = stream.detect_pcm16_stream(pcm, enc, cents, thr, window_seconds=3.0)
dt = time.perf_counter() - t0
prio = priority_ranks(species, api.PRIORITY_ORDER)
pred, best, okh, _ = eng.encode_detect_host(torch.from_numpy(pcm[: 256 * L].reshape(256, L)),
                                            np.stack([cents[s] for s in species]), np.array([thr[s] for s in species]), prio)
same = all((r.species == (species[p] if p >= 0 else None)) and r.best_distance == float(b) for r, p, b in zip(res, pred, best))
print(json.dumps({"workload": f"{HOURS:g} h of 48 kHz mono PCM_16 in {n_win} non-overlapping 3 s windows, one GPU",
                  "seconds": dt, "windows_per_s": n_win / dt, "times_real_time": HOURS * 3600 / dt,
                  "host_gbs": n_win * L * 2 / dt / 1e9, "detected": sum(r.detected for r in res),
                  "first_256_equal_chunk_api": bool(same)}))
