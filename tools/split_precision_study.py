#!/usr/bin/env python
"""CPU study: how many split-precision passes does the folded STFT GEMM need?  (no GPU; numpy emulation of the operands)

The CUDA path multiplies fp16 hi/lo splits of both operands in three tensor-core passes (A_hi B_hi + A_lo B_hi + A_hi B_lo,
DESIGN.md section 3).  Dropping a pass would remove a third of the issued flops, and dropping ``A_hi B_lo`` also the B_lo
loads (19 % of the operand bytes of dftf3_kernel).  This script rebuilds the operands exactly as the kernels define them --
normalised PCM_16 samples x periodic Hann window, folded three times in float32, scaled by a power of two
into fp16 range, DFT matrix x 2^10 -- quantises them to fp16 hi / lo, forms the products of each variant with exact
accumulation (so that only operand quantisation is measured), and pushes the result through |X|^2 -> slaney mel -> dB ->
z-score -> crop.  Reported: max |feature - reference| / max |reference| against the reference-made fixtures
(tests/golden/feat_*.npz; the GPU tests hold 2e-4), and the same for the latent means through the stand-in encoder (1e-3).

    python tools/split_precision_study.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

from oracle import hotpath as hp  # noqa: E402
from oracle import librosa_port as lp  # noqa: E402

N, HOP, SR = 2048, 384, 48000
GOLDEN = REPO / "tests" / "golden"


def frames_of(y: np.ndarray) -> np.ndarray:
    """Reflect-padded, Hann-windowed frames in float32: u[f, k] = w[k] * y_pad[f * hop + k]."""
    ypad = np.pad(y.astype(np.float32), N // 2, mode="reflect")
    F = 1 + y.shape[0] // HOP
    idx = np.arange(F)[:, None] * HOP + np.arange(N)[None]
    w = (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(N) / N)).astype(np.float32)
    return (ypad[idx] * w[None]).astype(np.float32)


def fold3(u: np.ndarray):
    """Three-level fold in float32 -> per class (cos operand, sin operand, edge term)."""
    f32 = np.float32
    Q, H = N // 4, N // 2
    E = np.zeros((u.shape[0], H + 1), f32)
    O = np.zeros((u.shape[0], H + 1), f32)
    E[:, 0], E[:, H] = u[:, 0], u[:, H]
    E[:, 1:H] = u[:, 1:H] + u[:, :H:-1]
    O[:, 1:H] = u[:, 1:H] - u[:, :H:-1]
    k = np.arange(Q)
    # odd bins: K = N/4
    odd_c = (E[:, k] - E[:, H - k]).astype(f32)
    odd_s = (O[:, k] + O[:, H - k]).astype(f32)
    odd_edge = O[:, Q].copy()                                  # * sin(pi b / 2) -> Im
    # even bins: P, R over k <= N/4, folded again about N/8
    P = np.zeros((u.shape[0], Q + 1), f32)
    R = np.zeros((u.shape[0], Q + 1), f32)
    P[:, :Q] = E[:, k] + E[:, H - k]
    P[:, Q] = E[:, Q]
    R[:, 1:Q] = O[:, 1:Q] - O[:, H - np.arange(1, Q)]
    j = np.arange(Q // 2)
    e0_c = (P[:, j] + P[:, Q - j]).astype(f32)                 # b = 0 mod 4
    e0_s = (R[:, j] - R[:, Q - j]).astype(f32)
    e0_edge = P[:, Q // 2].copy()                              # * cos(pi b / 4) -> Re
    e2_c = (P[:, j] - P[:, Q - j]).astype(f32)                 # b = 2 mod 4
    e2_s = (R[:, j] + R[:, Q - j]).astype(f32)
    e2_edge = R[:, Q // 2].copy()                              # * sin(pi b / 4) -> Im
    return {"odd": (odd_c, odd_s, odd_edge), "e0": (e0_c, e0_s, e0_edge), "e2": (e2_c, e2_s, e2_edge)}


def split16(x: np.ndarray):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64)


def spectrum(u: np.ndarray, variant: str, bins: np.ndarray) -> np.ndarray:
    """|X[b]|^2 for the listed bins from the folded operands; ``variant`` in exact | 3pass | noBlo | noAlo | 1pass."""
    ops = fold3(u)
    amax = max(float(np.abs(v[0]).max()) for v in ops.values())
    sa = 2.0 ** np.floor(np.log2(16384.0 / max(amax, 1e-30)))        # per-chunk power of two, operands within +-2^14
    sb = 1024.0
    out = np.zeros((u.shape[0], bins.shape[0]))
    for name, sel in (("odd", bins % 2 == 1), ("e0", bins % 4 == 0), ("e2", bins % 4 == 2)):
        b = bins[sel].astype(np.float64)
        c_op, s_op, edge = ops[name]
        K = c_op.shape[1]
        th = 2.0 * np.pi * np.outer(np.arange(K), b) / N
        Bc, Bs = np.cos(th), np.sin(th)
        if name == "odd":
            e_re, e_im = 0.0, np.outer(edge.astype(np.float64), np.sin(np.pi * b / 2.0))
        elif name == "e0":
            e_re, e_im = np.outer(edge.astype(np.float64), np.cos(np.pi * b / 4.0)), 0.0
        else:
            e_re, e_im = 0.0, np.outer(edge.astype(np.float64), np.sin(np.pi * b / 4.0))
        if variant == "exact":
            re, im = c_op.astype(np.float64) @ Bc, s_op.astype(np.float64) @ Bs
        else:
            def prod(a32, B):
                ah, al = split16((a32 * np.float32(sa)).astype(np.float32))
                bh, bl = split16((B * sb).astype(np.float32))
                r = ah @ bh
                if variant in ("3pass", "noBlo"):
                    r = r + al @ bh
                if variant in ("3pass", "noAlo"):
                    r = r + ah @ bl
                return r / (sa * sb)
            re, im = prod(c_op, Bc), prod(s_op, Bs)
        out[:, sel] = (re + e_re) ** 2 + (im + e_im) ** 2
    return out


def features_from_power(pw: np.ndarray, bins: np.ndarray, fb: np.ndarray) -> np.ndarray:
    S = (fb[:, bins].astype(np.float64) @ pw.T).astype(np.float32)          # [n_mels, F]
    S_db = lp.power_to_db(S, ref=np.max)
    S_db = (S_db - S_db.mean()) / (S_db.std() + 1e-8)
    return hp.crop_or_pad_time(S_db, target_frames=192).astype(np.float32)


def main():
    from conftest import load_pcm_case
    from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder
    enc = build_standin_encoder(seed=123)
    fb = lp.mel_filterbank(sr=SR, n_fft=N, n_mels=64, fmin=150.0, fmax=15000.0)
    bins = np.nonzero(fb.sum(axis=0) > 0)[0]
    variants = ("exact", "3pass", "noBlo", "noAlo", "1pass")
    worst = {v: [0.0, 0.0] for v in variants}
    print(f"{'case':14s} " + " ".join(f"{v + ' feat':>12s} {v + ' mu':>10s}" for v in variants))
    for key in ("noise_3s", "tonal_3s", "pulsed_3s", "burst0_3s", "burst3_3s", "hot_3s"):
        x, d = load_pcm_case(GOLDEN / f"feat_{key}.npz")
        y, _, _ = hp.rms_normalize_batch(x[None], pcm16=True)
        u = frames_of(y[0])
        ref_feat, ref_mu = d["feat"], d["z"]
        row = []
        for v in variants:
            feat = features_from_power(spectrum(u, v, bins), bins, fb)
            mu = hp.encode_features(enc, feat)
            ef = float(np.max(np.abs(feat - ref_feat)) / np.max(np.abs(ref_feat)))
            em = float(np.max(np.abs(mu - ref_mu)) / np.max(np.abs(ref_mu)))
            worst[v] = [max(worst[v][0], ef), max(worst[v][1], em)]
            row.append(f"{ef:12.2e} {em:10.2e}")
        print(f"{key:14s} " + " ".join(row))
    print(f"{'worst':14s} " + " ".join(f"{worst[v][0]:12.2e} {worst[v][1]:10.2e}" for v in variants))
    print("limits held by the GPU tests: features 2e-4, latents 1e-3")


if __name__ == "__main__":
    main()
