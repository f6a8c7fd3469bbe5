"""Stand-in VAE encoder + the layer-program exporter the CUDA encoder consumes.

The thesis encoder (``soundscape_vae`` ``_BirdNet`` factory + ``bird_net_vae_audio_splitted.yaml``)
is not part of the reference tree (docs/REPRODUCE_THESIS_BASELINE.md:36-42), so this module
supplies an architecture that satisfies every constraint the reference code places on it
(SURVEY.md appendix A):

* Hydra ``_target_`` resolves to a *factory*; calling the factory returns the ``nn.Module``
  (map_detector_core.py:135-147).
* input ``[B, 1, T=192, M=64]`` float32 (map_detector_core.py:267-268), eval mode
  (map_detector_core.py:179), at least one ``nn.Linear`` (07_encode_wav_to_latent.py:195-199).
* returns ``(mu, logvar)``; the reference takes the first tensor = the latent mean
  (map_detector_core.py:275-278).

``export_program`` walks *any* module built from Conv2d / BatchNorm2d / ReLU / MaxPool2d /
Flatten / Linear leaves (it does not hard-code this stand-in), folds eval-mode BatchNorm into the
preceding convolution, keeps only the layers on the path to the latent mean, and re-verifies the
exported program against the module before handing it to the CUDA library.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class BirdNetVAEEncoder(nn.Module):
    """BirdNET-flavoured conv stack -> (mu, logvar)."""

    def __init__(self, in_frames: int = 192, in_mels: int = 64,
                 channels: Sequence[int] = (32, 64, 128, 128), hidden: int = 512,
                 latent_dim: int = 128):
        super().__init__()
        layers: List[nn.Module] = []
        c_in, h, w = 1, in_frames, in_mels
        for c_out in channels:
            layers += [nn.Conv2d(c_in, c_out, kernel_size=3, stride=1, padding=1),
                       nn.BatchNorm2d(c_out), nn.ReLU(inplace=False), nn.MaxPool2d(2)]
            c_in, h, w = c_out, h // 2, w // 2
        self.features = nn.Sequential(*layers)
        self.flatten = nn.Flatten()
        self.fc = nn.Linear(c_in * h * w, hidden)
        self.act = nn.ReLU(inplace=False)
        self.fc_mu = nn.Linear(hidden, latent_dim)
        self.fc_logvar = nn.Linear(hidden, latent_dim)
        self.latent_dim = latent_dim

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        h = self.act(self.fc(self.flatten(self.features(x))))
        return self.fc_mu(h), self.fc_logvar(h)


class BirdNetVAEEncoderFactory:
    """Hydra ``_target_``: ``instantiate(cfg['encoder'])`` returns this; ``factory()`` -> module."""

    def __init__(self, **kwargs: Any):
        self.kwargs = kwargs

    def __call__(self) -> nn.Module:
        return BirdNetVAEEncoder(**self.kwargs)


def init_standin_weights(module: nn.Module, seed: int = 123) -> nn.Module:
    """Deterministic random init: He-normal convs/linears so activations
    stay O(1), and *randomised BatchNorm running statistics / affine* so that BN folding is tested."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                fan_in = m.weight[0].numel()
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (2.0 / fan_in) ** 0.5)
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
            elif isinstance(m, nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.4 + 0.8)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    return module.eval()


def build_standin_encoder(seed: int = 123, **kwargs: Any) -> nn.Module:
    return init_standin_weights(BirdNetVAEEncoderFactory(**kwargs)(), seed=seed)


# ----------------------------------------------------------------------------------------
# Layer program
# ----------------------------------------------------------------------------------------
@dataclass
class ConvOp:
    weight: np.ndarray            # [Cout, kh, kw, Cin] float32, BN folded (K-major: tap, then cin)
    bias: np.ndarray              # [Cout] float32
    stride: int
    pad: int
    relu: bool
    pool: int                     # 1 = none, 2 = fused MaxPool2d(2)
    in_hw: Tuple[int, int]        # input H, W
    out_hw: Tuple[int, int]       # output H, W after pooling


@dataclass
class LinearOp:
    weight: np.ndarray            # [out, in] float32; columns in NHWC-flatten order for the first
    bias: np.ndarray
    relu: bool


@dataclass
class EncoderProgram:
    in_hw: Tuple[int, int]
    ops: List[Any] = field(default_factory=list)
    latent_dim: int = 0

    def flops_per_chunk(self) -> float:
        total = 0.0
        for op in self.ops:
            if isinstance(op, ConvOp):
                cout, kh, kw, cin = op.weight.shape
                oh, ow = op.out_hw[0] * op.pool, op.out_hw[1] * op.pool
                total += 2.0 * cout * oh * ow * cin * kh * kw
            else:
                total += 2.0 * op.weight.shape[0] * op.weight.shape[1]
        return total


class UnsupportedEncoder(RuntimeError):
    pass


def _first_tensor(out: Any) -> torch.Tensor:
    """Which output tensor the reference treats as the latent (map_detector_core.py:272-290)."""
    if isinstance(out, torch.Tensor):
        return out
    if isinstance(out, (list, tuple)):
        t = next((z for z in out if isinstance(z, torch.Tensor)), None)
    elif isinstance(out, dict):
        t = None
        for k in ("z", "latent", "mu", "mean", "embedding"):
            if k in out and isinstance(out[k], torch.Tensor):
                t = out[k]
                break
        if t is None:
            t = next((v for v in out.values() if isinstance(v, torch.Tensor)), None)
    else:
        t = None
    if t is None:
        raise UnsupportedEncoder(f"cannot find a latent tensor in encoder output {type(out)}")
    return t


def export_program(module: nn.Module, in_frames: int = 192, in_mels: int = 64,
                   verify: bool = True) -> EncoderProgram:
    """Trace one forward with leaf hooks, keep the chain input -> latent mean, fuse
    Conv+BN+ReLU+MaxPool, permute the first Linear to NHWC flatten order."""
    module = module.eval()
    records: List[Tuple[nn.Module, int, torch.Tensor]] = []
    keep: List[torch.Tensor] = []

    def hook(m, inp, out):
        if not (len(inp) == 1 and isinstance(inp[0], torch.Tensor) and isinstance(out, torch.Tensor)):
            raise UnsupportedEncoder(f"leaf {type(m).__name__} is not tensor -> tensor")
        keep.extend([inp[0], out])
        records.append((m, id(inp[0]), out))

    handles = [m.register_forward_hook(hook) for m in module.modules() if not list(m.children())]
    try:
        with torch.no_grad():
            x = torch.randn(2, 1, in_frames, in_mels, generator=torch.Generator().manual_seed(7))
            keep.append(x)
            out = module(x)
    finally:
        for h in handles:
            h.remove()
    latent = _first_tensor(out)
    if latent.ndim != 2:
        raise UnsupportedEncoder(f"latent of rank {latent.ndim} (segment pooling) is not supported yet")

    # walk the execution record backwards from the latent (handles in-place leaves, where the
    # output tensor *is* the input tensor, because earlier producers are met later in the scan)
    chain: List[nn.Module] = []
    cur = id(latent)
    for m, iid, o in reversed(records):
        if cur == id(x):
            break
        if id(o) == cur:
            chain.append(m)
            cur = iid
    if cur != id(x):
        raise UnsupportedEncoder("a non-module op sits between leaf modules on the path to the "
                                 "latent mean (functional op / residual add); wrap it in nn.Module")
    chain.reverse()

    prog = EncoderProgram(in_hw=(in_frames, in_mels))
    h, w, c = in_frames, in_mels, 1
    flat_from: Optional[Tuple[int, int, int]] = None
    i = 0
    while i < len(chain):
        m = chain[i]
        if isinstance(m, nn.Conv2d):
            if flat_from is not None:
                raise UnsupportedEncoder("Conv2d after Flatten")
            if m.groups != 1 or m.dilation != (1, 1) or m.kernel_size[0] != m.kernel_size[1] \
                    or m.stride[0] != m.stride[1] or m.padding[0] != m.padding[1] or m.padding_mode != "zeros":
                raise UnsupportedEncoder(f"unsupported Conv2d configuration: {m}")
            wgt = m.weight.detach().double()
            b = m.bias.detach().double() if m.bias is not None else torch.zeros(m.out_channels, dtype=torch.float64)
            j = i + 1
            if j < len(chain) and isinstance(chain[j], nn.BatchNorm2d):
                bn = chain[j]
                s = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
                wgt = wgt * s[:, None, None, None]
                b = (b - bn.running_mean.detach().double()) * s + bn.bias.detach().double()
                j += 1
            relu = j < len(chain) and isinstance(chain[j], nn.ReLU)
            j += int(relu)
            pool = 1
            if j < len(chain) and isinstance(chain[j], nn.MaxPool2d):
                mp = chain[j]
                ks = mp.kernel_size if isinstance(mp.kernel_size, int) else mp.kernel_size[0]
                st = mp.stride if isinstance(mp.stride, int) else mp.stride[0]
                if ks != 2 or st != 2 or mp.padding not in (0, (0, 0)) or not relu:
                    raise UnsupportedEncoder(f"unsupported MaxPool2d: {mp}")
                pool = 2
                j += 1
            k, st, pd = m.kernel_size[0], m.stride[0], m.padding[0]
            oh = (h + 2 * pd - k) // st + 1
            ow = (w + 2 * pd - k) // st + 1
            if pool == 2 and (oh % 2 or ow % 2):
                raise UnsupportedEncoder("MaxPool2d(2) on odd spatial size")
            prog.ops.append(ConvOp(weight=wgt.permute(0, 2, 3, 1).contiguous().float().numpy(),
                                   bias=b.float().numpy(), stride=st, pad=pd, relu=relu, pool=pool,
                                   in_hw=(h, w), out_hw=(oh // pool, ow // pool)))
            h, w, c = oh // pool, ow // pool, m.out_channels
            i = j
        elif isinstance(m, nn.Flatten):
            flat_from = (c, h, w)
            i += 1
        elif isinstance(m, nn.Linear):
            first_linear = not any(isinstance(o, LinearOp) for o in prog.ops)
            if first_linear and flat_from is None and (h, w) != (1, 1):
                raise UnsupportedEncoder("Linear on an un-flattened feature map")
            wgt = m.weight.detach().float()
            if first_linear and flat_from is not None:
                cc, hh, ww = flat_from
                if wgt.shape[1] != cc * hh * ww:
                    raise UnsupportedEncoder("first Linear does not match the flattened feature map")
                wgt = wgt.view(-1, cc, hh, ww).permute(0, 2, 3, 1).reshape(wgt.shape[0], -1)
            b = m.bias.detach().float() if m.bias is not None else torch.zeros(m.out_features)
            relu = i + 1 < len(chain) and isinstance(chain[i + 1], nn.ReLU)
            prog.ops.append(LinearOp(weight=wgt.contiguous().numpy(), bias=b.numpy(), relu=relu))
            i += 1 + int(relu)
        elif isinstance(m, (nn.Dropout, nn.Dropout2d, nn.Identity)):
            i += 1  # identity in eval mode
        else:
            raise UnsupportedEncoder(f"unsupported layer type on the latent path: {type(m).__name__}")
    if not prog.ops or not isinstance(prog.ops[-1], LinearOp):
        raise UnsupportedEncoder("the latent mean must be produced by an nn.Linear")
    prog.latent_dim = int(prog.ops[-1].weight.shape[0])

    if verify:
        with torch.no_grad():
            ref = _first_tensor(module(x))
            got = run_program_torch(prog, x)
        err = float((ref - got).abs().max() / ref.abs().max().clamp_min(1e-12))
        if err > 1e-4:
            raise UnsupportedEncoder(f"exported program does not reproduce the module (rel err {err:.3e})")
    return prog


def run_program_torch(prog: EncoderProgram, x: torch.Tensor) -> torch.Tensor:
    """fp32 torch replay of an exported program on ``x [B,1,T,M]`` (export self-check and the
    fp32 reference the CUDA encoder is compared with in tests)."""
    h = x.permute(0, 2, 3, 1).contiguous()                    # NHWC
    for op in prog.ops:
        if isinstance(op, ConvOp):
            wt = torch.from_numpy(op.weight).permute(0, 3, 1, 2).contiguous()
            y = F.conv2d(h.permute(0, 3, 1, 2), wt, torch.from_numpy(op.bias), stride=op.stride, padding=op.pad)
            if op.relu:
                y = F.relu(y)
            if op.pool == 2:
                y = F.max_pool2d(y, 2)
            h = y.permute(0, 2, 3, 1).contiguous()
        else:
            h = h.reshape(h.shape[0], -1)
            h = F.linear(h, torch.from_numpy(op.weight), torch.from_numpy(op.bias))
            if op.relu:
                h = F.relu(h)
    return h
