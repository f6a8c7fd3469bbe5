"""CPU: `export_program` (torch.fx walk of the encoder module) reproduces the module in fp32 for every layer type SURVEY
appendix A lists, interprets the output the way the reference does (map_detector_core.py:272-295), and refuses loudly what
the CUDA encoder cannot run."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from amphibian_vae_latent_detector_b200.encoder import (AddOp, AffineOp, ConvOp, LinearOp, PoolOp, UnsupportedEncoder,
                                                        _first_tensor, build_residual_standin_encoder, build_standin_encoder,
                                                        export_program, init_standin_weights, reduce_latent, run_program_torch)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def test_chain_standin_exports_as_before():
    prog = export_program(build_standin_encoder(123))
    assert [type(o).__name__ for o in prog.ops] == ["ConvOp"] * 4 + ["LinearOp"] * 2
    assert prog.latent_dim == 128 and prog.n_seg == 1 and prog.out_nchw is None
    assert all(o.pool == 2 and not o.pool_avg and o.relu for o in prog.ops[:4])


def test_residual_segmented_standin():
    mod = build_residual_standin_encoder()
    prog = export_program(mod, 384, 64)
    assert prog.n_seg == 2 and prog.in_hw == (192, 64) and prog.latent_dim == 128
    kinds = [type(o) for o in prog.ops]
    # the four residual adds ride in the epilogue of the later of their two producing convolutions (no AddOp left)
    assert kinds.count(AddOp) == 0 and sum(isinstance(o, ConvOp) and o.residual >= 0 for o in prog.ops) == 4
    assert any(isinstance(o, ConvOp) and o.stride == 2 and o.weight.shape[1] == 3 for o in prog.ops)
    assert any(isinstance(o, ConvOp) and o.weight.shape[1] == 1 and o.stride == 2 for o in prog.ops)        # 1x1 stride-2 shortcut
    assert any(isinstance(o, PoolOp) and o.k == 0 for o in prog.ops) and any(isinstance(o, PoolOp) and o.k == 2 and o.avg for o in prog.ops)
    x = torch.randn(3, 1, 384, 64, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        out = mod(x)
        assert out["mu"].shape == (3, 2, 128)                     # rank 3: the reference averages over dim 1
        ref = reduce_latent(_first_tensor(out))
    assert _rel(run_program_torch(prog, x), ref) < 1e-5
    # logvar is not on the path to the latent: its head is not exported
    assert sum(isinstance(o, LinearOp) for o in prog.ops) == 2


def test_an_add_that_cannot_be_fused_stays_an_op():
    class TwoBranch(nn.Module):                                   # both operands are activated: neither convolution can absorb the add
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

    mod = TwoBranch().eval()
    prog = export_program(mod, 192, 64)
    assert sum(isinstance(o, AddOp) for o in prog.ops) == 1 and all(o.residual < 0 for o in prog.ops if isinstance(o, ConvOp))
    x = torch.randn(2, 1, 192, 64, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        assert _rel(run_program_torch(prog, x), mod(x)) < 1e-5


def test_output_conventions_follow_the_reference():
    class Head(nn.Module):
        def __init__(self, mode):
            super().__init__()
            self.mode = mode
            self.f = nn.Sequential(nn.Conv2d(1, 32, 3, 1, 1), nn.ReLU(), nn.MaxPool2d(2), nn.Flatten(), nn.Linear(32 * 96 * 32, 16))
            self.g = nn.Linear(16, 8)

        def forward(self, x):
            z = self.f(x)
            if self.mode == "dict_pref":
                return {"aux": self.g(z), "latent": z}            # "latent" wins over dict order (core:281-285)
            if self.mode == "dict_first":
                return {"a": self.g(z), "b": z}                   # no known key: the first tensor value
            if self.mode == "list":
                return [None, self.g(z), z]                       # first tensor in the sequence
            return z

    for mode, dim in (("dict_pref", 16), ("dict_first", 8), ("list", 8), ("plain", 16)):
        mod = init_standin_weights(Head(mode), seed=2)
        prog = export_program(mod)
        assert prog.latent_dim == dim, mode


def test_unsupported_modules_are_refused():
    class Grouped(nn.Module):
        def __init__(self):
            super().__init__()
            self.c = nn.Conv2d(1, 8, 3, 1, 1)
            self.d = nn.Conv2d(8, 8, 3, 1, 1, groups=8)
            self.fc = nn.Linear(8 * 192 * 64, 4)

        def forward(self, x):
            return self.fc(self.d(self.c(x)).flatten(1))

    with pytest.raises(UnsupportedEncoder, match="Conv2d configuration"):
        export_program(init_standin_weights(Grouped(), 1))

    class Gelu(nn.Module):
        def __init__(self):
            super().__init__()
            self.c = nn.Conv2d(1, 8, 3, 1, 1)
            self.fc = nn.Linear(8 * 192 * 64, 4)

        def forward(self, x):
            return self.fc(torch.nn.functional.gelu(self.c(x)).flatten(1))

    with pytest.raises(UnsupportedEncoder, match="unsupported operation"):
        export_program(init_standin_weights(Gelu(), 1))

    class Branchy(nn.Module):                                       # data-dependent control flow: fx cannot trace it
        def __init__(self):
            super().__init__()
            self.fc = nn.Linear(192 * 64, 4)

        def forward(self, x):
            if x.sum() > 0:
                x = x * 2
            return self.fc(x.flatten(1))

    with pytest.raises(UnsupportedEncoder):
        export_program(Branchy())
