#!/usr/bin/env python
"""sha256 of the features / latents of a fixed set of synthetic chunks through every operand source -- run once per build
(AVLD_LIB_PATH selects the library) to show that a kernel change left the bits alone:
    python tools/feature_hash.py; AVLD_LIB_PATH=.../libavld_prev.so python tools/feature_hash.py"""
import hashlib, json, os, sys
from pathlib import Path
import torch
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from amphibian_vae_latent_detector_b200 import synth
from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder
from amphibian_vae_latent_detector_b200.engine import Engine


def h(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]


out = {"lib": os.environ.get("AVLD_LIB_PATH", "libavld.so")}
enc = build_standin_encoder(seed=123)
for L, tag in ((144000, "3s"), (240000, "5s"), (48000, "1s")):
    x, _ = synth.make_chunks(96, L, seed=11, special_every=7)
    x = x.cuda()
    eng = Engine(0, chunk_len=L, max_batch=64)
    eng.load_encoder(enc)
    f_q, _, _ = eng.normalize_logmel(x, pcm16=True)         # prep integers -> fold3<2>
    f_f, _, _ = eng.normalize_logmel(x, pcm16=False)        # float chunk -> fold3<0>, no PCM_16 round trip
    y, _, _ = eng.rms_normalize(x, pcm16=False)
    f_y = eng.logmel(y)                                     # already normalised input
    pcm = torch.clamp(torch.round(x * 32767.0), -32768, 32767).to(torch.int16)
    mu_p, _ = eng.encode(pcm)                               # raw PCM_16 chunk
    mu_f, _ = eng.encode(x)
    out[tag] = {"feat_q16": h(f_q), "feat_float": h(f_f), "feat_logmel": h(f_y), "mu_pcm": h(mu_p), "mu_float": h(mu_f),
                "nan": bool(torch.isnan(f_q).any())}
    eng.close()
print(json.dumps(out))
