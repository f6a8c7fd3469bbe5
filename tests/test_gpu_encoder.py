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


def _check_against_torch(mod, feat, *, target_frames=192, max_batch=8, tol=1e-4):
    from amphibian_vae_latent_detector_b200.encoder import reduce_latent, _first_tensor
    from amphibian_vae_latent_detector_b200.engine import Engine
    eng = Engine(0, chunk_len=144000, max_batch=max_batch, target_frames=target_frames)
    prog = eng.load_encoder(mod)
    with torch.no_grad():
        ref = reduce_latent(_first_tensor(mod(feat[:, None]))).numpy()   # what the reference makes of the output (core:272-295)
    mu = eng.encoder_forward(feat.cuda()).cpu().numpy()
    eng.close()
    assert mu.shape == ref.shape, (mu.shape, ref.shape)
    assert rel(mu, ref) < tol, rel(mu, ref)
    return prog


def test_residual_segmented_encoder_vs_torch_module():
    """configs/bird_net_res_vae_audio_splitted.yaml: residual adds, stride-2 and 1x1 convolutions, channel counts that are
    not tcgen05 tile sizes (48, 96), AvgPool2d, a global average pool, dict output with a rank-3 latent [B, n_seg, C]."""
    import yaml
    from pathlib import Path
    from amphibian_vae_latent_detector_b200 import reference_api as api
    from amphibian_vae_latent_detector_b200.encoder import init_standin_weights, ConvOp, PoolOp
    cfg = yaml.safe_load((Path(__file__).resolve().parents[1] / "configs" / "bird_net_res_vae_audio_splitted.yaml").read_text())
    mod = init_standin_weights(api.build_nn_module(api._instantiate(api.pick_encoder_cfg(cfg))), seed=321)
    g = torch.Generator().manual_seed(5)
    prog = _check_against_torch(mod, torch.randn(19, 192, 64, generator=g))
    # the residual adds are fused into the epilogue of the convolution that produces their later operand
    assert prog.n_seg == 1 and sum(isinstance(o, ConvOp) and o.residual >= 0 for o in prog.ops) == 4
    assert any(isinstance(o, PoolOp) and o.k == 0 for o in prog.ops)
    # two segments per chunk: the latent is the mean of the segment latents (core:292-293)
    prog2 = _check_against_torch(mod, torch.randn(9, 384, 64, generator=g), target_frames=384)
    assert prog2.n_seg == 2


def test_encoder_layer_variety_vs_torch_module():
    """Layer types one at a time: flatten head after residual stages, average pooling fused behind a convolution, a
    BatchNorm that cannot be folded (pre-activation), 5x5 stride-2 stem, stand-alone max pooling, a feature-map latent
    (flattened in NCHW order by the reference), tuple output."""
    import torch.nn as nn
    from amphibian_vae_latent_detector_b200.encoder import build_residual_standin_encoder, init_standin_weights

    g = torch.Generator().manual_seed(11)
    feat = torch.randn(6, 192, 64, generator=g)
    _check_against_torch(build_residual_standin_encoder(head="flatten", widths=(64, 128), blocks=(1, 1)), feat)

    class PreAct(nn.Module):
        def __init__(self):
            super().__init__()
            self.stem = nn.Conv2d(1, 24, 5, 2, 2)                       # 5x5 stride-2 stem, 24 channels (padded to 32)
            self.bn0 = nn.BatchNorm2d(24)
            self.c1 = nn.Conv2d(24, 64, 3, 1, 1)
            self.pool1 = nn.AvgPool2d(2)                                # fused behind the convolution + ReLU
            self.bn1 = nn.BatchNorm2d(64)                               # after the pooling: cannot be folded
            self.c2 = nn.Conv2d(64, 64, 3, 1, 1, bias=False)
            self.mp = nn.MaxPool2d(3, 2)                                # stand-alone, overlapping windows
            self.c3 = nn.Conv2d(64, 40, 1)                              # 1x1
            self.fc = nn.Linear(40 * 23 * 7, 96)

        def forward(self, x):
            h = torch.relu(self.bn0(self.stem(x)))
            h = self.pool1(torch.relu(self.c1(h)))
            h = self.c2(torch.relu(self.bn1(h))) + h
            h = self.c3(self.mp(h))
            return self.fc(torch.flatten(h, 1)), h                      # tuple: the first tensor is the latent

    _check_against_torch(init_standin_weights(PreAct(), seed=3), feat)

    class MapLatent(nn.Module):                                         # rank-4 output: flattened [B, C*H*W] (core:294-295)
        def __init__(self):
            super().__init__()
            self.f = nn.Sequential(nn.Conv2d(1, 32, 3, 1, 1), nn.ReLU(), nn.MaxPool2d(2), nn.Conv2d(32, 64, 3, 2, 1), nn.ReLU(),
                                   nn.AvgPool2d(4), nn.Conv2d(64, 20, 1))

        def forward(self, x):
            return {"embedding": self.f(x)}

    _check_against_torch(init_standin_weights(MapLatent(), seed=4), feat)

    class TwoBranch(nn.Module):                                         # an add of two activated branches: stays a kernel of its own
        def __init__(self):
            super().__init__()
            self.stem = nn.Conv2d(1, 32, 3, 1, 1)
            self.a = nn.Conv2d(32, 64, 3, 1, 1)
            self.b = nn.Conv2d(32, 64, 1)
            self.fc = nn.Linear(64, 16)

        def forward(self, x):
            h = torch.nn.functional.max_pool2d(torch.relu(self.stem(x)), 2)
            h = torch.relu(self.a(h)) + torch.relu(self.b(h))
            return self.fc(h.mean(dim=(2, 3)))

    from amphibian_vae_latent_detector_b200.encoder import AddOp
    prog = _check_against_torch(init_standin_weights(TwoBranch(), seed=6), feat)
    assert any(isinstance(o, AddOp) for o in prog.ops)
