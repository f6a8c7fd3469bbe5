#!/usr/bin/env python
"""A tiny pass over every kernel of the library for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
12 chunks through rms_normalize, the fused encode (chain stand-in) and a residual / segmented stand-in, the radial fit
over the q_out grid, the decision, the Gaussian-MAP fit + scoring, the resampler and the host-buffer call."""
import sys
from pathlib import Path
import numpy as np
import torch
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from amphibian_vae_latent_detector_b200 import synth
from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder, build_residual_standin_encoder
from amphibian_vae_latent_detector_b200.engine import Engine, priority_ranks

L, n = 144000, 12
dev = torch.device("cuda", 0)
x, lab = synth.make_chunks(n, L, seed=5, device=dev)
species = ["Batrachyla_leptopus", "Batrachyla_taeniata", "Calyptocephalella_gayi", "Pleurodema_thaul"]
prio = priority_ranks(species, species)
for build in (build_standin_encoder, build_residual_standin_encoder):
    eng = Engine(0, chunk_len=L, max_batch=8)
    eng.load_encoder(build())
    y, ok, rms = eng.rms_normalize(x)
    Z, ok = eng.encode(x, pcm16=True)
    fit = eng.fit_radial(Z, lab, 4, 0.95, (0.10, 0.25))
    pred, best = eng.decide(fit.radii_local, torch.from_numpy(fit.rk[0]).to(dev), torch.from_numpy(prio).to(dev))
    mfit = eng.fit_map(Z, lab, species, cov_type="lda", cov_structure="diag", shrink=0.1)
    mp = eng.map_score(Z, mfit)
    xh = torch.clamp(torch.round(x * 32767.0), -32768, 32767).to(torch.int16).cpu().pin_memory()
    out = eng.encode_detect_host(xh, np.nan_to_num(fit.centroids), fit.rk[0], prio, pcm16=True)
    r = eng.resample(x[0, :44100].contiguous(), 44100, 48000)
    torch.cuda.synchronize()
    print(build.__name__, "pred", pred.tolist(), "host", [int(v) for v in out[0]][:12], "resampled", tuple(r.shape))
    eng.close()
print("sanitizer workload done")
