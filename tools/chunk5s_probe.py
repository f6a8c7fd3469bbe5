#!/usr/bin/env python
"""Secondary chunk length (chunk_seconds = 5.0, the default of 08 / 09 / 10: 08:392-396): resident and host-buffer
throughput on one GPU, same step definition as bench.py but L = 240 000 (F = 626 frames)."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from amphibian_vae_latent_detector_b200 import synth  # noqa: E402
from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder  # noqa: E402
from amphibian_vae_latent_detector_b200.engine import Engine, priority_ranks  # noqa: E402

L, N, MB = 240000, 8192, 512
species = ["Batrachyla_leptopus", "Batrachyla_taeniata", "Calyptocephalella_gayi", "Pleurodema_thaul"]
eng = Engine(0, chunk_len=L, max_batch=MB)
eng.load_encoder(build_standin_encoder(seed=123))
X = torch.empty(N, L, device="cuda")
lab = torch.empty(N, dtype=torch.int32, device="cuda")
for i in range(0, N, 512):
    xs, ls = synth.make_chunks(512, L, seed=123, first_index=i, device="cuda")
    X[i:i + 512], lab[i:i + 512] = xs, ls
prio = priority_ranks(species, species)


def step():
    Z, ok = eng.encode(X, pcm16=True)
    fit = eng.fit_radial(Z, lab, 4, 0.95, (0.10, 0.15, 0.20, 0.25))
    pred, best = eng.decide(fit.radii_local, torch.from_numpy(fit.rk[0]).cuda(), torch.from_numpy(prio).cuda())
    return fit, torch.bincount((pred + 1).long(), minlength=5).cpu()


for _ in range(2):
    fit, hist = step()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    fit, hist = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
xh = torch.empty(N, L, dtype=torch.int16, pin_memory=True)
for i in range(0, N, 1024):
    xh[i:i + 1024].copy_(torch.clamp(torch.round(X[i:i + 1024] * 32767.0), -32768, 32767).to(torch.int16))
torch.cuda.synchronize()
cent, thr = np.nan_to_num(fit.centroids), fit.rk[0]
eng.encode_detect_host(xh, cent, thr, prio)
t0 = time.perf_counter()
for _ in range(3):
    eng.encode_detect_host(xh, cent, thr, prio)
dt = (time.perf_counter() - t0) / 3
print(json.dumps({"workload": f"{N} synthetic 5 s chunks (L = {L}), one GPU, max_batch {MB}", "dft_mode": eng.dft_info()["mode"],
                  "value_chunks_per_s": N / (ms / 1e3), "audio_seconds_per_s": 5.0 * N / (ms / 1e3),
                  "e2e_chunks_per_s": N / dt, "e2e_h2d_gbs": N * L * 2 / dt / 1e9, "decision_hist": hist.tolist()}))
