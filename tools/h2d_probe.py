#!/usr/bin/env python
"""Where does the host-buffer path's time go?  (a) raw pinned H2D bandwidth of this box, alone and while the encode
kernels run; (b) avld_encode_detect_host_pcm16 throughput against the slab size (max_batch); (c) the per-slab
device timeline (AVLD_HOST_TRACE).  Writes gpurun_out/h2d_probe.json and gpurun_out/host_trace_*.txt."""
import json
import os
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

L = 144000
out = {}
dev = torch.device("cuda", 0)
N = int(os.environ.get("PROBE_CHUNKS", "8192"))
x, lab = synth.make_chunks(1024, L, seed=123, device=dev)
x16d = torch.clamp(torch.round(x * 32767.0), -32768, 32767).to(torch.int16)
xh = torch.empty(N, L, dtype=torch.int16, pin_memory=True)
for i in range(0, N, 1024):
    xh[i:i + 1024].copy_(x16d[: min(1024, N - i)])
torch.cuda.synchronize()


def h2d_bw(nbytes_rows, reps=8, busy=None):
    dst = torch.empty(nbytes_rows, L, dtype=torch.int16, device=dev)
    s = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        dst.copy_(xh[:nbytes_rows], non_blocking=True)
    torch.cuda.synchronize()
    if busy is not None:
        busy()
    with torch.cuda.stream(s):
        e0.record()
        for r in range(reps):
            dst.copy_(xh[(r * nbytes_rows) % (N - nbytes_rows + 1):][:nbytes_rows], non_blocking=True)
        e1.record()
    torch.cuda.synchronize()
    return nbytes_rows * L * 2 * reps / (e0.elapsed_time(e1) / 1e3) / 1e9


for rows in (64, 256, 1024):
    out[f"h2d_gbs_alone_{rows}rows"] = h2d_bw(rows)

species = ["Batrachyla_leptopus", "Batrachyla_taeniata", "Calyptocephalella_gayi", "Pleurodema_thaul"]
prio = priority_ranks(species, species)
enc = build_standin_encoder(seed=123)
for mb in (256, 512, 1024, 2048):
    eng = Engine(0, chunk_len=L, max_batch=mb)
    eng.load_encoder(enc)
    Z, ok = eng.encode(x, pcm16=True)
    fit = eng.fit_radial(Z, lab, 4, 0.95, [0.25])
    cent, thr = np.nan_to_num(fit.centroids), fit.rk[0]
    if mb == 1024:
        X8 = x.repeat(4, 1)

        def busy():
            eng.encode(X8, pcm16=True)       # ~4096 chunks of kernels queued on the default stream
        out["h2d_gbs_under_compute_1024rows"] = h2d_bw(1024, reps=6, busy=busy)
        torch.cuda.synchronize()
        # device-resident throughput for comparison
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.encode(X8, pcm16=True)
        e1.record()
        torch.cuda.synchronize()
        out["resident_chunks_per_s"] = X8.shape[0] / (e0.elapsed_time(e1) / 1e3)
        del X8
    for _ in range(2):
        eng.encode_detect_host(xh, cent, thr, prio, pcm16=True)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        eng.encode_detect_host(xh, cent, thr, prio, pcm16=True)
    dt = time.perf_counter() - t0
    out[f"e2e_chunks_per_s_mb{mb}"] = N * reps / dt
    os.environ["AVLD_HOST_TRACE"] = str(REPO / "gpurun_out" / f"host_trace_mb{mb}.txt")
    eng.encode_detect_host(xh, cent, thr, prio, pcm16=True)
    del os.environ["AVLD_HOST_TRACE"]
    eng.close()
    del eng
    print(json.dumps(out), flush=True)

(REPO / "gpurun_out").mkdir(exist_ok=True)
(REPO / "gpurun_out" / "h2d_probe.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out, indent=1))
