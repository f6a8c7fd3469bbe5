#!/usr/bin/env python
"""CPU study: operand precision the encoder's tensor-core layers need (no GPU; torch float64 emulation).

The CUDA encoder multiplies bf16 hi / lo splits of activations and weights in three passes.  This
script replays the exported layer program with the operands of every tensor-core layer (all convs but the first, which
runs in fp32 on the CUDA cores, and the dense layers) quantised as a variant prescribes, exact accumulation, and reports the
latent error ``max|a - b| / max|b|`` against the unquantised float64 replay on the reference-made feature fixtures, plus the
largest activation the variant has to represent (fp16 overflows at 65504).

    python tools/encoder_precision_study.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

from amphibian_vae_latent_detector_b200.encoder import ConvOp, build_standin_encoder, export_program  # noqa: E402

GOLDEN = REPO / "tests" / "golden"
VARIANTS = ("exact", "bf16x3", "bf16x1", "fp16x1", "fp16x2_act", "fp16x2_wgt", "fp16x3")


def split(x: torch.Tensor, dt):
    hi = x.to(dt)
    lo = (x.to(torch.float32) - hi.to(torch.float32)).to(dt)
    return hi.to(torch.float64), lo.to(torch.float64)


def contract(fn, a: torch.Tensor, w: torch.Tensor, variant: str) -> torch.Tensor:
    """``fn(a, w)`` (a conv or a matmul, float64) with the operands quantised as ``variant`` says."""
    if variant == "exact":
        return fn(a.double(), w.double())
    dt = torch.bfloat16 if variant.startswith("bf16") else torch.float16
    a_hi, a_lo = split(a.float(), dt)
    w_hi, w_lo = split(w.float(), dt)
    out = fn(a_hi, w_hi)
    if variant.endswith("x3") or variant.endswith("x2_act"):
        out = out + fn(a_lo, w_hi)
    if variant.endswith("x3") or variant.endswith("x2_wgt"):
        out = out + fn(a_hi, w_lo)
    return out


def replay(prog, x: torch.Tensor, variant: str):
    h = x.permute(0, 2, 3, 1).contiguous().double()
    amax = 0.0
    first = True
    for op in prog.ops:
        if isinstance(op, ConvOp):
            wt = torch.from_numpy(op.weight).permute(0, 3, 1, 2).contiguous()
            conv = lambda a, w, op=op: F.conv2d(a, w, None, stride=op.stride, padding=op.pad)   # noqa: E731
            y = contract(conv, h.permute(0, 3, 1, 2), wt, "exact" if first else variant)          # conv1: fp32 CUDA cores
            y = y + torch.from_numpy(op.bias).double()[None, :, None, None]
            if op.relu:
                y = F.relu(y)
            if op.pool == 2:
                y = F.max_pool2d(y, 2)
            h = y.permute(0, 2, 3, 1).contiguous().float().double()                               # stored as fp32-exact hi + lo
            first = False
        else:
            h = h.reshape(h.shape[0], -1)
            y = contract(lambda a, w: a @ w.T, h, torch.from_numpy(op.weight), variant) + torch.from_numpy(op.bias).double()
            h = (F.relu(y) if op.relu else y).float().double()
        amax = max(amax, float(h.abs().max()))
    return h.numpy(), amax


def main():
    enc = build_standin_encoder(seed=123)
    prog = export_program(enc)
    keys = ("noise_3s", "tonal_3s", "pulsed_3s", "burst0_3s", "burst3_3s", "hot_3s", "silent_3s", "pulsed_5s", "tonal_5s")
    feats = np.stack([np.load(GOLDEN / f"feat_{k}.npz")["feat"].T for k in keys])               # [n, 192, 64]
    x = torch.from_numpy(feats)[:, None]
    with torch.no_grad():
        ref, amax = replay(prog, x, "exact")
        print(f"largest activation anywhere: {amax:.1f}; latents max |mu| {np.abs(ref).max():.2f}")
        for v in VARIANTS[1:]:
            mu, _ = replay(prog, x, v)
            per = np.max(np.abs(mu - ref), axis=1) / np.max(np.abs(ref), axis=1)
            print(f"{v:11s} worst latent error {per.max():.2e}   median {np.median(per):.2e}   (contract: 1e-3)")


if __name__ == "__main__":
    main()
