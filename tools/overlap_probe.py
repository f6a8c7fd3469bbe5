#!/usr/bin/env python
"""Do the HBM-bound and the tensor-bound kernels of consecutive passes overlap when the passes alternate between two
contexts on two streams?  One context / one stream (what avld_encode does) against two and three lanes:
    python tools/overlap_probe.py [chunks] [reps]"""
import json, sys
from pathlib import Path
import torch
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from amphibian_vae_latent_detector_b200 import synth
from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder
from amphibian_vae_latent_detector_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
B = 1024
x, _ = synth.make_chunks(n, 144000, seed=5, special_every=5)
x = x.cuda()
enc = build_standin_encoder(seed=123)
out = {}
ref = None
for lanes in (1, 2, 3):
    engs = [Engine(0, chunk_len=144000, max_batch=B) for _ in range(lanes)]
    for e in engs:
        e.load_encoder(enc)
    streams = [torch.cuda.Stream() for _ in range(lanes)]
    mu = torch.empty(n, engs[0].latent_dim, device="cuda")

    def run():
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        for i, lo in enumerate(range(0, n, B)):
            k = i % lanes
            with torch.cuda.stream(streams[k]):
                m, _ok = engs[k].encode(x[lo:lo + B])
                mu[lo:lo + B].copy_(m)
        for s in streams:
            cur.wait_stream(s)

    for _ in range(2):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if ref is None:
        ref = mu.clone()
    out[f"lanes_{lanes}"] = {"chunks_per_s": round(reps * n / (ms / 1e3)), "ms_per_1024_chunks": round(ms / (reps * n / B), 3),
                             "bit_identical_to_one_lane": bool(torch.equal(mu, ref))}
    for e in engs:
        e.close()
    del engs
    torch.cuda.empty_cache()
print(json.dumps(out))
