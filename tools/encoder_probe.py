#!/usr/bin/env python
"""Encoder-only throughput of the two stand-ins (chain and residual / segmented) with the per-stage CUDA-event timers:
    python tools/encoder_probe.py [chunks]"""
import json, sys
from pathlib import Path
import torch
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder, build_residual_standin_encoder
from amphibian_vae_latent_detector_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
feat = torch.randn(n, 192, 64, device="cuda")
for name, build in (("chain", build_standin_encoder), ("residual", build_residual_standin_encoder)):
    eng = Engine(0, chunk_len=144000, max_batch=1024)
    prog = eng.load_encoder(build())
    for _ in range(2):
        eng.encoder_forward(feat)
    eng.collect(reset=True); eng.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(3):
        eng.encoder_forward(feat)
    e1.record(); torch.cuda.synchronize()
    st = eng.collect(reset=True)
    passes = 3 * (n // 1024)
    print(json.dumps({"encoder": name, "ops": len(prog.ops), "gflop_per_chunk": round(prog.flops_per_chunk() / 1e9, 3) if hasattr(prog, "flops_per_chunk") else None,
                      "ms_per_1024_chunks": round(e0.elapsed_time(e1) / passes, 3), "chunks_per_s": round(3 * n / (e0.elapsed_time(e1) / 1e3)),
                      "stage_ms_per_pass": {k: round(v["ms"] / passes, 3) for k, v in st.items() if v["timed_launches"]},
                      "launches_per_pass": {k: v["timed_launches"] // passes for k, v in st.items() if v["timed_launches"]}}))
    eng.close()
