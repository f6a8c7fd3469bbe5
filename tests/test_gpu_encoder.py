"""GPU: the encoder forward (map_detector_core.py:270-300) on tcgen05 vs the fp32 torch module."""
import numpy as np
import pytest
import torch

from oracle import hotpath as hp

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-12))


def test_encoder_vs_torch_module(engine3s, standin_encoder):
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(70, 192, 64, generator=g)           # z-scored features are O(1)
    with torch.no_grad():
        ref = standin_encoder(feat[:, None])[0].numpy()    # first tensor of (mu, logvar) = the latent mean
    mu = engine3s.encoder_forward(feat.cuda()).cpu().numpy()
    assert mu.shape == ref.shape == (70, 128)
    assert rel(mu, ref) < 1e-4, rel(mu, ref)
    # batch-1 execution, exactly like the reference's per-file loop
    one = hp.encode_features(standin_encoder, feat[5].numpy().T)
    assert rel(mu[5], one) < 1e-4


def test_encoder_other_architecture():
    """The CUDA encoder is driven by the exported layer program, not hard-wired to the stand-in."""
    from amphibian_vae_latent_detector_b200.encoder import BirdNetVAEEncoder, init_standin_weights
    from amphibian_vae_latent_detector_b200.engine import Engine
    mod = init_standin_weights(BirdNetVAEEncoder(channels=(32, 64, 64), hidden=256, latent_dim=64), seed=9)
    eng = Engine(0, chunk_len=144000, max_batch=8)
    eng.load_encoder(mod)
    feat = torch.randn(11, 192, 64, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = mod(feat[:, None])[0].numpy()
    mu = eng.encoder_forward(feat.cuda()).cpu().numpy()
    assert rel(mu, ref) < 1e-4
    eng.close()
