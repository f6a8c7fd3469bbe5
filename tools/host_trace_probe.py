#!/usr/bin/env python
"""Per-slab device timeline of avld_encode_detect_host_pcm16 (AVLD_HOST_TRACE, read at context creation) for calls of 4096
and 16384 chunks, plus the wall time of repeated calls.  Writes gpurun_out/host_trace_<n>.txt."""
import os, sys, time, json
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
(REPO / "gpurun_out").mkdir(exist_ok=True)
os.environ["AVLD_HOST_TRACE"] = str(REPO / "gpurun_out" / "host_trace.txt")
import numpy as np, torch
from amphibian_vae_latent_detector_b200 import synth
from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder
from amphibian_vae_latent_detector_b200.engine import Engine, priority_ranks
L, dev = 144000, torch.device("cuda", 0)
N = 16384
x, lab = synth.make_chunks(1024, L, seed=123, device=dev)
x16d = torch.clamp(torch.round(x * 32767.0), -32768, 32767).to(torch.int16)
xh = torch.empty(N, L, dtype=torch.int16, pin_memory=True)
for i in range(0, N, 1024):
    xh[i:i + 1024].copy_(x16d)
torch.cuda.synchronize()
species = ["a", "b", "c", "d"]
prio = priority_ranks(species, species)
eng = Engine(0, chunk_len=L, max_batch=1024)
eng.load_encoder(build_standin_encoder(seed=123))
Z, ok = eng.encode(x, pcm16=True)
fit = eng.fit_radial(Z, lab, 4, 0.95, [0.25])
cent, thr = np.nan_to_num(fit.centroids), fit.rk[0]
out = {}
for n in (4096, 16384):
    for _ in range(2):
        eng.encode_detect_host(xh[:n], cent, thr, prio, pcm16=True)
    os.replace(REPO / "gpurun_out" / "host_trace.txt", REPO / "gpurun_out" / f"host_trace_{n}.txt")
    t0 = time.perf_counter()
    reps = 4
    for _ in range(reps):
        eng.encode_detect_host(xh[:n], cent, thr, prio, pcm16=True)
    dt = (time.perf_counter() - t0) / reps
    out[n] = {"ms_per_call": dt * 1e3, "chunks_per_s": n / dt, "gbs": n * L * 2 / dt / 1e9}
print(json.dumps(out))
print(open(REPO / "gpurun_out" / "host_trace_4096.txt").read())
