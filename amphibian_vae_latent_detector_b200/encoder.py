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
# A second stand-in with everything SURVEY appendix A lists: residual blocks, stride-2 and 1x1
# convolutions, average pooling, a global average pool, a time axis split into segments and a
# dict output whose latent is [B, n_seg, C]
# ----------------------------------------------------------------------------------------
class _ResBlock(nn.Module):
    def __init__(self, c_in: int, c_out: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(c_in, c_out, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(c_out)
        self.conv2 = nn.Conv2d(c_out, c_out, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(c_out)
        self.shortcut = None
        if stride != 1 or c_in != c_out:
            self.shortcut = nn.Sequential(nn.Conv2d(c_in, c_out, 1, stride, 0, bias=False), nn.BatchNorm2d(c_out))

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        r = x if self.shortcut is None else self.shortcut(x)
        return torch.relu(y + r)


class BirdNetResVAEEncoder(nn.Module):
    """BirdNET-style residual encoder over 192-frame segments: ``[B,1,T,M] -> {"mu": [B, T/192, D], "logvar": ...}``.
    The reference averages a rank-3 latent over its segment axis (map_detector_core.py:292-293)."""

    def __init__(self, seg_frames: int = 192, in_mels: int = 64, stem: int = 32, widths: Sequence[int] = (48, 96, 128),
                 blocks: Sequence[int] = (1, 2, 1), latent_dim: int = 128, head: str = "gap"):
        super().__init__()
        self.seg_frames, self.latent_dim, self.head = seg_frames, latent_dim, head
        self.stem = nn.Sequential(nn.Conv2d(1, stem, 3, 1, 1, bias=False), nn.BatchNorm2d(stem), nn.ReLU(), nn.MaxPool2d(2))
        layers: List[nn.Module] = []
        c = stem
        for i, (wd, nb) in enumerate(zip(widths, blocks)):
            for b in range(nb):
                layers.append(_ResBlock(c, wd, 2 if (b == 0 and i > 0) else 1))
                c = wd
        self.stages = nn.Sequential(*layers)
        self.pool = nn.AvgPool2d(2)
        h, w = seg_frames // 2, in_mels // 2
        for i in range(1, len(widths)):
            h, w = (h + 1) // 2, (w + 1) // 2
        h, w = h // 2, w // 2
        self.gap = nn.AdaptiveAvgPool2d(1)
        feat = c if head == "gap" else c * h * w
        self.fc = nn.Linear(feat, 256)
        self.fc_mu = nn.Linear(256, latent_dim)
        self.fc_logvar = nn.Linear(256, latent_dim)

    def forward(self, x):
        b, _, t, m = x.shape
        x = x.reshape(b * (t // self.seg_frames), 1, self.seg_frames, m)
        h = self.pool(self.stages(self.stem(x)))
        h = self.gap(h).flatten(1) if self.head == "gap" else h.flatten(1)
        h = torch.relu(self.fc(h))
        return {"mu": self.fc_mu(h).view(b, -1, self.latent_dim), "logvar": self.fc_logvar(h).view(b, -1, self.latent_dim)}


class BirdNetResVAEEncoderFactory:
    """Hydra ``_target_`` of configs/bird_net_res_vae_audio_splitted.yaml."""

    def __init__(self, **kwargs: Any):
        self.kwargs = kwargs

    def __call__(self) -> nn.Module:
        return BirdNetResVAEEncoder(**self.kwargs)


def build_residual_standin_encoder(seed: int = 321, **kwargs: Any) -> nn.Module:
    return init_standin_weights(BirdNetResVAEEncoderFactory(**kwargs)(), seed=seed)


# ----------------------------------------------------------------------------------------
# Layer program: a small dataflow graph over numbered tensors (tensor 0 = one feature segment
# [1, seg_frames, n_mels]).  Images are NCHW here and NHWC on the device.
# ----------------------------------------------------------------------------------------
@dataclass
class ConvOp:
    weight: np.ndarray            # [Cout, kh, kw, Cin] float32, BN folded (K-major: tap, then cin)
    bias: np.ndarray              # [Cout] float32
    stride: int
    pad: int
    relu: bool
    pool: int                     # 1 = none, 2 = fused 2x2 / stride-2 pooling after the activation
    in_hw: Tuple[int, int]        # input H, W
    out_hw: Tuple[int, int]       # output H, W after pooling
    pool_avg: bool = False        # the fused pooling is AvgPool2d(2) instead of MaxPool2d(2)
    src: int = -1
    dst: int = -1
    residual: int = -1            # tensor added to the convolution's output before the activation (a fused residual add)


@dataclass
class LinearOp:
    weight: np.ndarray            # [out, in] float32; columns in NHWC-flatten order when the input is an image
    bias: np.ndarray
    relu: bool
    src: int = -1
    dst: int = -1


@dataclass
class AddOp:                      # dst = a + b (+ ReLU): the residual connection
    a: int
    b: int
    dst: int
    relu: bool


@dataclass
class AffineOp:                   # dst = src * scale[c] + shift[c] (+ ReLU): a BatchNorm that could not be folded, a lone ReLU
    scale: np.ndarray
    shift: np.ndarray
    relu: bool
    src: int
    dst: int


@dataclass
class PoolOp:                     # stand-alone Max / AvgPool2d(k, stride) (no padding), or the global average (k = 0)
    avg: bool
    k: int
    stride: int
    in_hw: Tuple[int, int]
    out_hw: Tuple[int, int]
    src: int
    dst: int


@dataclass
class EncoderProgram:
    in_hw: Tuple[int, int]                      # (seg_frames, n_mels) of one segment image
    ops: List[Any] = field(default_factory=list)
    latent_dim: int = 0
    n_seg: int = 1                              # segments per chunk: the latent is the mean over them (core:292-293)
    out: int = -1                               # tensor id of the latent
    out_nchw: Optional[Tuple[int, int, int]] = None   # the latent is a feature map (C, H, W), flattened in NCHW order (core:294-295)
    shapes: dict = field(default_factory=dict)  # tensor id -> (C, H, W) image or (D,) vector

    def flops_per_chunk(self) -> float:
        total = 0.0
        for op in self.ops:
            if isinstance(op, ConvOp):
                cout, kh, kw, cin = op.weight.shape
                oh, ow = op.out_hw[0] * op.pool, op.out_hw[1] * op.pool
                total += 2.0 * cout * oh * ow * cin * kh * kw
            elif isinstance(op, LinearOp):
                total += 2.0 * op.weight.shape[0] * op.weight.shape[1]
        return total * self.n_seg


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


def reduce_latent(t: torch.Tensor) -> torch.Tensor:
    """core:292-295: rank 3 -> mean over dim 1, higher ranks flattened."""
    if t.ndim == 3:
        t = t.mean(dim=1)
    if t.ndim > 2:
        t = t.reshape(t.shape[0], -1)
    return t


# ---- graph capture -------------------------------------------------------------------------------------------
_RELU_FNS = (torch.relu, F.relu, torch.nn.functional.relu)


class _Val:
    """What a traced node evaluates to while the program is being built: a tensor tagged with its role."""
    __slots__ = ("t", "kind", "tid", "pending")

    def __init__(self, t, kind, tid):
        self.t, self.kind, self.tid = t, kind, tid      # kind: "img" [N,C,H,W] | "vec" [N,D] | "tok" [B,n_seg,D]


def _as_int(v):
    return v if isinstance(v, int) else (v[0] if isinstance(v, (tuple, list)) else int(v))


def export_program(module: nn.Module, in_frames: int = 192, in_mels: int = 64, verify: bool = True) -> EncoderProgram:
    """``torch.fx`` trace of the module -> dataflow program.  Supported on the path to the latent: Conv2d (square kernel,
    stride 1 / 2, zero padding, no groups / dilation), eval-mode BatchNorm2d (folded into the convolution in front of it when
    it is that convolution's only consumer), ReLU, MaxPool2d / AvgPool2d without padding, AdaptiveAvgPool2d(1) / mean over
    (H, W), residual ``+``, Flatten / view, Linear, Dropout / Identity, a leading reshape that splits the time axis into
    equal segments and the regrouping view behind the head.  The exported program is replayed in fp32 against the module;
    anything it cannot express raises :class:`UnsupportedEncoder`."""
    import operator
    import torch.fx as fx

    module = module.eval()
    try:
        gm = fx.symbolic_trace(module)
    except Exception as exc:                                   # data-dependent control flow etc.
        raise UnsupportedEncoder(f"torch.fx cannot trace the encoder: {exc}") from exc
    B = 2
    x = torch.randn(B, 1, in_frames, in_mels, generator=torch.Generator().manual_seed(7))
    mods = dict(gm.named_modules())
    users = {n: list(n.users) for n in gm.graph.nodes}

    prog = EncoderProgram(in_hw=(in_frames, in_mels))
    shapes: dict = {}
    next_id = [1]

    def new_tensor(shape) -> int:
        tid = next_id[0]
        next_id[0] += 1
        shapes[tid] = tuple(int(v) for v in shape)
        return tid

    env: dict = {}

    def val(a):
        if isinstance(a, fx.Node):
            return env[a]
        if isinstance(a, (tuple, list)):
            return type(a)(val(v) for v in a)
        if isinstance(a, dict):
            return {k: val(v) for k, v in a.items()}
        return a

    def raw(v):                                                 # the plain Python / tensor value behind a traced value
        if isinstance(v, _Val):
            return v.t
        if isinstance(v, (tuple, list)):
            return type(v)(raw(u) for u in v)
        if isinstance(v, dict):
            return {k: raw(u) for k, u in v.items()}
        return v

    def img(t, tid):
        return _Val(t, "img", tid)

    def relu_of(v: _Val, t: torch.Tensor) -> _Val:
        """ReLU applied to value v: folded into the producing conv / linear / add / affine when v has no other consumer."""
        last = prog.ops[-1] if prog.ops else None
        if last is not None and getattr(last, "dst", None) == v.tid and not getattr(last, "relu", True) \
                and not (isinstance(last, ConvOp) and last.pool != 1) and v.pending == 1:
            last.relu = True
            return _Val(t, v.kind, v.tid)
        c = t.shape[1]
        tid = new_tensor(shapes[v.tid])
        prog.ops.append(AffineOp(np.ones(c, np.float32), np.zeros(c, np.float32), True, v.tid, tid))
        return _Val(t, v.kind, tid)

    def pool_of(v: _Val, t: torch.Tensor, avg: bool, k: int, st: int, pad) -> _Val:
        if v.kind != "img" or pad not in (0, (0, 0)):
            raise UnsupportedEncoder("pooling with padding / on a non-image tensor")
        h, w = v.t.shape[2], v.t.shape[3]
        last = prog.ops[-1] if prog.ops else None
        if isinstance(last, ConvOp) and last.dst == v.tid and last.pool == 1 and v.pending == 1 and k == 2 and st == 2 \
                and h % 2 == 0 and w % 2 == 0 and (last.relu or not avg or True):
            last.pool, last.pool_avg = 2, avg
            last.out_hw = (h // 2, w // 2)
            shapes[v.tid] = (t.shape[1], t.shape[2], t.shape[3])
            return img(t, v.tid)
        tid = new_tensor((t.shape[1], t.shape[2], t.shape[3]))
        prog.ops.append(PoolOp(avg, k, st, (h, w), (t.shape[2], t.shape[3]), v.tid, tid))
        return img(t, tid)

    def gap_of(v: _Val, t4: torch.Tensor) -> _Val:
        tid = new_tensor((t4.shape[1], 1, 1))
        prog.ops.append(PoolOp(True, 0, 1, (v.t.shape[2], v.t.shape[3]), (1, 1), v.tid, tid))
        return img(t4, tid)

    def reshape_of(v: _Val, t: torch.Tensor) -> _Val:
        """view / reshape / flatten / squeeze of a traced tensor, recognised by its shapes."""
        src = v.t
        if v.kind == "img" and t.ndim == 2 and t.shape[0] == src.shape[0]:          # [N,C,H,W] -> [N, C*H*W]: Flatten
            return _Val(t, "vec", v.tid)
        if v.kind == "img" and t.ndim == 4 and t.shape == src.shape:
            return _Val(t, "img", v.tid)
        if v.kind == "vec" and t.ndim == 3 and t.shape[0] == B and t.shape[0] * t.shape[1] == src.shape[0] \
                and t.shape[2] == src.shape[1] and t.shape[1] == prog.n_seg:          # [B*n_seg, D] -> [B, n_seg, D]
            return _Val(t, "tok", v.tid)
        if v.kind == "vec" and t.ndim == 2 and t.shape == src.shape:
            return _Val(t, "vec", v.tid)
        if v.kind == "tok" and t.ndim == 2 and prog.n_seg == 1 and t.shape[0] == B:
            return _Val(t, "vec", v.tid)
        raise UnsupportedEncoder(f"reshape {tuple(src.shape)} -> {tuple(t.shape)} of a {v.kind} tensor is not supported")

    with torch.no_grad():
        for node in gm.graph.nodes:
            if node.op == "placeholder":
                env[node] = _Val(x, "input", 0)
                env[node].pending = len(users[node])
                continue
            if node.op == "get_attr":
                raise UnsupportedEncoder(f"free tensor attribute {node.target} on the latent path")
            if node.op == "output":
                env[node] = val(node.args[0])
                continue
            args, kwargs = val(node.args), val(node.kwargs)
            tvals = [a for a in list(args) + list(kwargs.values()) if isinstance(a, _Val)]
            # plain Python arithmetic on sizes etc.: just evaluate
            if node.op == "call_module":
                m = mods[node.target]
                result = m(*raw(args), **raw(kwargs))
            elif node.op == "call_function":
                result = node.target(*raw(args), **raw(kwargs))
            else:
                result = getattr(raw(args[0]), node.target)(*raw(args[1:]), **raw(kwargs))
            if not isinstance(result, torch.Tensor) or not tvals:
                env[node] = result
                continue
            v = tvals[0]
            out_v: Optional[_Val] = None
            tgt = node.target
            mod = mods[tgt] if node.op == "call_module" else None
            name = tgt if isinstance(tgt, str) else getattr(tgt, "__name__", str(tgt))

            if v.kind == "input":
                # the only thing allowed on the raw input: the split of the time axis into equal segments (or nothing)
                if mod is None and name in ("reshape", "view", "contiguous", "unsqueeze", "squeeze", "float", "to"):
                    if result.ndim == 4 and result.shape[1] == 1 and result.shape[3] == in_mels \
                            and result.shape[0] * result.shape[2] == B * in_frames and in_frames % result.shape[2] == 0 \
                            and torch.equal(result.reshape(B, 1, in_frames, in_mels), x):
                        prog.n_seg = in_frames // result.shape[2]
                        prog.in_hw = (int(result.shape[2]), in_mels)
                        shapes[0] = (1, int(result.shape[2]), in_mels)
                        out_v = _Val(result, "img", 0)
                    else:
                        raise UnsupportedEncoder(f"unsupported reshape of the input to {tuple(result.shape)}")
                else:
                    shapes.setdefault(0, (1, in_frames, in_mels))
                    v = _Val(x, "img", 0)
                    v.pending = len(users[node.args[0]]) if isinstance(node.args[0], fx.Node) else 1
            if out_v is None:
                shapes.setdefault(0, (1, in_frames, in_mels))
                if isinstance(mod, nn.Conv2d):
                    if v.kind != "img":
                        raise UnsupportedEncoder("Conv2d on a non-image tensor")
                    if mod.groups != 1 or mod.dilation != (1, 1) or mod.kernel_size[0] != mod.kernel_size[1] \
                            or mod.stride[0] != mod.stride[1] or mod.padding[0] != mod.padding[1] or mod.padding_mode != "zeros" \
                            or isinstance(mod.padding, str) or mod.stride[0] not in (1, 2):
                        raise UnsupportedEncoder(f"unsupported Conv2d configuration: {mod}")
                    wgt = mod.weight.detach().double()
                    b = mod.bias.detach().double() if mod.bias is not None else torch.zeros(mod.out_channels, dtype=torch.float64)
                    tid = new_tensor((result.shape[1], result.shape[2], result.shape[3]))
                    prog.ops.append(ConvOp(weight=wgt, bias=b, stride=mod.stride[0], pad=mod.padding[0], relu=False, pool=1,
                                           in_hw=(v.t.shape[2], v.t.shape[3]), out_hw=(result.shape[2], result.shape[3]),
                                           src=v.tid, dst=tid))
                    out_v = img(result, tid)
                elif isinstance(mod, nn.BatchNorm2d):
                    s = mod.weight.detach().double() / torch.sqrt(mod.running_var.detach().double() + mod.eps)
                    sh = mod.bias.detach().double() - mod.running_mean.detach().double() * s
                    last = prog.ops[-1] if prog.ops else None
                    if isinstance(last, ConvOp) and last.dst == v.tid and not last.relu and last.pool == 1 and v.pending == 1:
                        last.weight = last.weight * s[:, None, None, None]
                        last.bias = last.bias * s + sh
                        out_v = img(result, v.tid)
                    else:
                        tid = new_tensor(shapes[v.tid])
                        prog.ops.append(AffineOp(s.float().numpy(), sh.float().numpy(), False, v.tid, tid))
                        out_v = img(result, tid)
                elif isinstance(mod, nn.ReLU) or (mod is None and (tgt in _RELU_FNS or name in ("relu", "relu_"))):
                    out_v = relu_of(v, result)
                elif isinstance(mod, (nn.MaxPool2d, nn.AvgPool2d)):
                    k = _as_int(mod.kernel_size)
                    st = _as_int(mod.stride if mod.stride is not None else mod.kernel_size)
                    if getattr(mod, "ceil_mode", False) or (isinstance(mod, nn.AvgPool2d) and not mod.count_include_pad and mod.padding not in (0, (0, 0))):
                        raise UnsupportedEncoder(f"unsupported pooling: {mod}")
                    out_v = pool_of(v, result, isinstance(mod, nn.AvgPool2d), k, st, mod.padding)
                elif mod is None and tgt in (F.max_pool2d, F.avg_pool2d):
                    k = _as_int(raw(args[1]) if len(args) > 1 else raw(kwargs["kernel_size"]))
                    st_raw = raw(args[2]) if len(args) > 2 else raw(kwargs.get("stride", None))
                    st = k if st_raw in (None, []) else _as_int(st_raw)
                    pad = raw(args[3]) if len(args) > 3 else raw(kwargs.get("padding", 0))
                    out_v = pool_of(v, result, tgt is F.avg_pool2d, k, st, pad)
                elif isinstance(mod, nn.AdaptiveAvgPool2d) or (mod is None and tgt is F.adaptive_avg_pool2d):
                    if tuple(result.shape[2:]) != (1, 1):
                        raise UnsupportedEncoder("AdaptiveAvgPool2d to a size other than 1")
                    out_v = gap_of(v, result)
                elif mod is None and name == "mean":
                    dims = raw(args[1]) if len(args) > 1 else raw(kwargs.get("dim"))
                    dims = tuple(dims) if isinstance(dims, (tuple, list)) else (dims,)
                    if v.kind == "img" and sorted(d % 4 for d in dims) == [2, 3]:
                        keep = bool(raw(kwargs.get("keepdim", False)) or (len(args) > 2 and raw(args[2])))
                        g = gap_of(v, result if keep else result[:, :, None, None])
                        out_v = g if keep else _Val(result, "vec", g.tid)
                    elif v.kind == "tok" and dims in ((1,), (-2,)):
                        out_v = _Val(result, "segmean", v.tid)
                    else:
                        raise UnsupportedEncoder(f"mean over dims {dims} of a {v.kind} tensor")
                elif mod is None and tgt in (operator.add, torch.add, operator.iadd) or (mod is None and name in ("add", "add_")):
                    if len(tvals) != 2 or tvals[0].t.shape != tvals[1].t.shape or tvals[0].kind != tvals[1].kind:
                        raise UnsupportedEncoder("`+` of a tensor with a scalar / of different shapes")
                    tid = new_tensor(shapes[tvals[0].tid])
                    prog.ops.append(AddOp(tvals[0].tid, tvals[1].tid, tid, False))
                    out_v = _Val(result, tvals[0].kind, tid)
                elif isinstance(mod, nn.Flatten) or (mod is None and (tgt is torch.flatten or name in (
                        "flatten", "view", "reshape", "squeeze", "unsqueeze", "contiguous"))):
                    out_v = reshape_of(v, result)
                elif isinstance(mod, nn.Linear):
                    if v.kind != "vec":
                        raise UnsupportedEncoder("Linear on an un-flattened tensor")
                    wgt = mod.weight.detach().float()
                    shp = shapes[v.tid]
                    if len(shp) == 3:                              # image flattened in NCHW order -> NHWC columns
                        cc, hh, ww = shp
                        if wgt.shape[1] != cc * hh * ww:
                            raise UnsupportedEncoder("Linear does not match the flattened feature map")
                        wgt = wgt.view(-1, cc, hh, ww).permute(0, 2, 3, 1).reshape(wgt.shape[0], -1)
                    b = mod.bias.detach().float() if mod.bias is not None else torch.zeros(mod.out_features)
                    tid = new_tensor((result.shape[1],))
                    prog.ops.append(LinearOp(weight=wgt.contiguous().numpy(), bias=b.numpy(), relu=False, src=v.tid, dst=tid))
                    out_v = _Val(result, "vec", tid)
                elif isinstance(mod, (nn.Dropout, nn.Dropout2d, nn.Identity)) or (mod is None and name in ("dropout", "float", "to", "clone", "detach")):
                    out_v = _Val(result, v.kind, v.tid)
                else:
                    raise UnsupportedEncoder(f"unsupported operation on the latent path: {node.op} {name}"
                                             f"{' (' + type(mod).__name__ + ')' if mod is not None else ''}")
            out_v.pending = len(users[node])
            env[node] = out_v

    out_val = env[[n for n in gm.graph.nodes if n.op == "output"][0]]

    def first_val(o):
        if isinstance(o, _Val):
            return o
        if isinstance(o, (list, tuple)):
            return next((z for z in o if isinstance(z, _Val)), None)
        if isinstance(o, dict):
            for k in ("z", "latent", "mu", "mean", "embedding"):
                if isinstance(o.get(k), _Val):
                    return o[k]
            return next((z for z in o.values() if isinstance(z, _Val)), None)
        return None

    lat = first_val(out_val)
    if lat is None:
        raise UnsupportedEncoder(f"cannot find a latent tensor in encoder output {type(out_val)}")
    if lat.kind == "input":
        raise UnsupportedEncoder("the encoder returns its input")
    if lat.kind == "tok" or lat.kind == "segmean":
        pass                                                       # [B, n_seg, D]: averaged over the segments (core:292-293)
    elif lat.kind == "vec":
        if prog.n_seg != 1:
            raise UnsupportedEncoder("a segmented encoder must return [B, n_seg, D]")
    elif lat.kind == "img":
        if prog.n_seg != 1:
            raise UnsupportedEncoder("a segmented encoder must return [B, n_seg, D]")
        prog.out_nchw = tuple(int(v) for v in shapes[lat.tid])     # flattened in NCHW order (core:294-295)
    prog.out = lat.tid

    # keep only what the latent needs, renumber nothing (tensor ids are just names)
    needed = {prog.out}
    kept = []
    for op in reversed(prog.ops):
        if op.dst in needed:
            kept.append(op)
            needed.update([op.a, op.b] if isinstance(op, AddOp) else [op.src])
    prog.ops = _fuse_residual_adds(list(reversed(kept)), prog.out)
    for op in prog.ops:
        if isinstance(op, ConvOp):
            op.weight = op.weight.permute(0, 2, 3, 1).contiguous().float().numpy()
            op.bias = op.bias.float().numpy()
            if op.src == 0 and op.weight.shape[3] != 1:
                raise UnsupportedEncoder("the first convolution must take the single-channel feature image")
    if not prog.ops:
        raise UnsupportedEncoder("no supported layer on the path to the latent")
    if any(isinstance(op, (AddOp, AffineOp, PoolOp, LinearOp)) and 0 in ([op.a, op.b] if isinstance(op, AddOp) else [op.src])
           for op in prog.ops):
        raise UnsupportedEncoder("the feature image must enter the network through a Conv2d")
    prog.shapes = {k: v for k, v in shapes.items() if k == 0 or any(k in (getattr(o, "dst", None), getattr(o, "src", None), getattr(o, "a", None), getattr(o, "b", None), getattr(o, "residual", None)) for o in prog.ops)}
    osh = shapes[prog.out]
    prog.latent_dim = int(np.prod(osh))

    if verify:
        with torch.no_grad():
            ref = reduce_latent(_first_tensor(module(x)))
            got = run_program_torch(prog, x)
        err = float((ref - got).abs().max() / ref.abs().max().clamp_min(1e-12))
        if ref.shape != got.shape or err > 1e-4:
            raise UnsupportedEncoder(f"exported program does not reproduce the module (shapes {tuple(ref.shape)} / "
                                     f"{tuple(got.shape)}, rel err {err:.3e})")
    return prog


def _inputs_of(op) -> list:
    if isinstance(op, AddOp):
        return [op.a, op.b]
    return [op.src] + ([op.residual] if isinstance(op, ConvOp) and op.residual >= 0 else [])


def _fuse_residual_adds(ops: list, out_tid: int) -> list:
    """`y = conv(x) + r` (a residual block's tail: BatchNorm already folded, ReLU after the sum): the add becomes the
    convolution's epilogue when the convolution is the LATER of the two producers (so r exists when it runs), has no
    activation / pooling of its own and feeds nothing else.  Saves one HBM round trip of the feature map per block."""
    ops = list(ops)
    changed = True
    while changed:
        changed = False
        uses: dict = {out_tid: 1}
        for op in ops:
            for t in _inputs_of(op):
                uses[t] = uses.get(t, 0) + 1
        producer = {op.dst: i for i, op in enumerate(ops)}
        for i, op in enumerate(ops):
            if not isinstance(op, AddOp):
                continue
            pa, pb = producer.get(op.a, -1), producer.get(op.b, -1)
            cand, other = (pa, op.b) if pa > pb else (pb, op.a)
            if cand < 0 or other == 0:
                continue
            p = ops[cand]
            if isinstance(p, ConvOp) and not p.relu and p.pool == 1 and p.residual < 0 and p.src != 0 and uses.get(p.dst, 0) == 1 \
                    and p.dst != other:
                p.residual, p.relu, p.dst = other, op.relu, op.dst
                del ops[i]
                changed = True
                break
    return ops


def run_program_torch(prog: EncoderProgram, x: torch.Tensor) -> torch.Tensor:
    """fp32 torch replay of an exported program on ``x [B,1,T,M]`` -> ``[B, latent_dim]`` (export self-check and the fp32
    reference the CUDA encoder is compared with in tests)."""
    b = x.shape[0]
    seg, mels = prog.in_hw
    t: dict = {0: x.reshape(b * prog.n_seg, 1, seg, mels)}
    for op in prog.ops:
        if isinstance(op, ConvOp):
            wt = torch.from_numpy(op.weight).permute(0, 3, 1, 2).contiguous()
            y = F.conv2d(t[op.src], wt, torch.from_numpy(op.bias), stride=op.stride, padding=op.pad)
            if op.residual >= 0:
                y = y + t[op.residual]
            if op.relu:
                y = F.relu(y)
            if op.pool == 2:
                y = F.avg_pool2d(y, 2) if op.pool_avg else F.max_pool2d(y, 2)
            t[op.dst] = y
        elif isinstance(op, LinearOp):
            h = t[op.src]
            if h.ndim == 4:
                h = h.permute(0, 2, 3, 1).reshape(h.shape[0], -1)      # NHWC flatten: the exported column order
            y = F.linear(h, torch.from_numpy(op.weight), torch.from_numpy(op.bias))
            t[op.dst] = F.relu(y) if op.relu else y
        elif isinstance(op, AddOp):
            y = t[op.a] + t[op.b]
            t[op.dst] = F.relu(y) if op.relu else y
        elif isinstance(op, AffineOp):
            h = t[op.src]
            shape = (1, -1, 1, 1) if h.ndim == 4 else (1, -1)
            y = h * torch.from_numpy(op.scale).view(shape) + torch.from_numpy(op.shift).view(shape)
            t[op.dst] = F.relu(y) if op.relu else y
        elif isinstance(op, PoolOp):
            h = t[op.src]
            if op.k == 0:
                t[op.dst] = h.mean(dim=(2, 3), keepdim=True)
            else:
                t[op.dst] = F.avg_pool2d(h, op.k, op.stride) if op.avg else F.max_pool2d(h, op.k, op.stride)
        else:
            raise TypeError(type(op))
    z = t[prog.out]
    z = z.reshape(z.shape[0], -1)                                      # NCHW flatten for a feature-map latent
    return z.reshape(b, prog.n_seg, -1).mean(dim=1)
