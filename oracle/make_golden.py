"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` from the reference's own code.

Run in the build container (needs ``/root/reference``)::

    python -m oracle.make_golden

The reference ships no golden vectors or tests, so the fixtures are the
outputs of the *unmodified* reference functions -- ``rms_normalize`` (00:29-38), ``wav_to_mel`` /
``encode_wav_to_latent`` (map_detector_core.py:198-300), ``fit_species_with_fp_control``
(08:310-333), ``get_detector_from_config`` (09:113-149), ``DetectorSession.predict_one``
(10:152-199) -- executed through ``oracle/ref_import.py`` on seeded synthetic chunks written as
PCM_16 WAV files, with the stand-in encoder.  Inputs are stored as int16 PCM (x = s/32768 * gain is
exact in float32) so that the fixtures are self-contained and small.
"""
from __future__ import annotations

import hashlib
import json
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

from oracle import librosa_port as lp  # noqa: E402
from oracle import ref_import  # noqa: E402
from amphibian_vae_latent_detector_b200 import synth  # noqa: E402
from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder  # noqa: E402

GOLDEN = REPO / "tests" / "golden"
MEL_KW = dict(sr=48000, n_mels=64, fmin=150.0, fmax=15000.0, hop_length=384, n_fft=2048, target_frames=192)
SPECIES = ["Batrachyla_leptopus", "Batrachyla_taeniata", "Calyptocephalella_gayi", "Pleurodema_thaul"]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def signal_bank(length: int, seed: int):
    """name -> float32 signal, all exactly representable as int16/32768 * gain."""
    x, _ = synth.make_chunks(8, length, seed=seed, special_every=0)
    x = x.numpy()
    rng = np.random.default_rng(seed)
    t = np.arange(length, dtype=np.float64) / 48000.0
    bank = {
        "noise": 0.03 * rng.standard_normal(length),
        "tonal": 0.2 * np.sin(2 * np.pi * 2600.0 * t) + 1e-4 * rng.standard_normal(length),
        "pulsed": x[1],
        "burst0": x[0],
        "burst3": x[3],
        "hot": 0.9 * np.sign(np.sin(2 * np.pi * 700.0 * t)) * (0.6 + 0.4 * rng.random(length)),
        "silent": 2e-5 * rng.standard_normal(length),
    }
    out = {}
    for k, v in bank.items():
        gain = np.float32(1.0)
        s = np.clip(np.rint(np.asarray(v, dtype=np.float64) * 32767.0), -32768, 32767).astype(np.int16)
        if k == "silent":
            gain = np.float32(2.0 ** -10)      # int16 alone cannot express rms < 1e-4 with detail
            s = np.clip(np.rint(np.asarray(v, dtype=np.float64) * 32767.0 * 1024.0), -32768, 32767).astype(np.int16)
        out[k] = (s, gain)
    return out


def to_float(s: np.ndarray, gain) -> np.ndarray:
    return (s.astype(np.float32) * np.float32(1.0 / 32768.0)) * np.float32(gain)


def main() -> None:
    if not ref_import.available():
        raise SystemExit("reference tree not available; fixtures can only be made in the build container")
    ref00, core, ref08, ref09, ref10 = (ref_import.load(k) for k in ("00", "core", "08", "09", "10"))
    GOLDEN.mkdir(parents=True, exist_ok=True)
    encoder = build_standin_encoder(seed=123)
    meta = {"numpy": np.__version__, "torch": torch.__version__, "mel_kw": MEL_KW, "species": SPECIES,
            "encoder": "amphibian_vae_latent_detector_b200.encoder.build_standin_encoder(seed=123)"}

    # ---------------- A. rms_normalize (00:29-38) ----------------
    rms_cases = {}
    for length, seed in ((144000, 11), (240000, 12), (4097, 13), (1000, 14), (100, 15), (7, 16)):
        for name, (s, gain) in signal_bank(length, seed).items():
            if length < 144000 and name not in ("noise", "hot", "silent"):
                continue
            x = to_float(s, gain)
            y, ok = ref00.rms_normalize(x)
            key = f"{name}_{length}"
            rms_cases[key] = dict(ok=bool(ok), sha=sha(np.asarray(y, dtype=np.float32)),
                                  rms=float(np.sqrt(np.mean(x ** 2))), head=np.asarray(y[:8], dtype=np.float32).tolist())
            np.savez_compressed(GOLDEN / f"rms_{key}.npz", pcm=s, gain=gain,
                                y_head=np.asarray(y[:64], dtype=np.float32), y_tail=np.asarray(y[-64:], dtype=np.float32),
                                y_stride=np.asarray(y[::997], dtype=np.float32))
    meta["rms_cases"] = rms_cases

    # ---------------- B/C. wav_to_mel + encode_wav_to_latent (core:198-300) via WAV files ----------------
    feat_cases = {}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        plan = []
        for name in ("noise", "tonal", "pulsed", "burst0", "burst3", "hot", "silent"):
            plan.append((f"{name}_3s", 144000, 21, name, 3.0, 144000))
        plan.append(("pulsed_5s", 240000, 22, "pulsed", 5.0, 240000))
        plan.append(("tonal_5s", 240000, 22, "tonal", 5.0, 240000))
        plan.append(("noise_short_pad", 100000, 23, "noise", 3.0, 144000))       # file shorter than duration -> zero pad
        plan.append(("burst0_long_trunc", 240000, 24, "burst0", 3.0, 144000))    # file longer -> truncate
        plan.append(("pulsed_1s_Tpad", 144000, 25, "pulsed", 1.0, 48000))        # F=126 < 192 -> frame padding
        for key, length, seed, name, duration, eff_len in plan:
            s, gain = signal_bank(length, seed)[name]
            x = to_float(s, gain)
            # the reference dataflow: 00 normalises + sf.write (PCM_16), then 08/09/10 load the file
            y_norm, ok = ref00.rms_normalize(x)
            wav = td / f"{key}.wav"
            sys.modules["soundfile"].write(wav, np.asarray(y_norm, dtype=np.float32), 48000)
            mel = core.wav_to_mel(wav, duration=duration, **MEL_KW).numpy()
            z = core.encode_wav_to_latent(encoder, wav, torch.device("cpu"), duration=duration, **MEL_KW)
            np.savez_compressed(GOLDEN / f"feat_{key}.npz", pcm=s, gain=gain, duration=np.float64(duration),
                                ok=np.uint8(ok), feat=mel.astype(np.float32), z=z.astype(np.float32))
            feat_cases[key] = dict(length=length, duration=duration, ok=bool(ok))

        # ---------------- E. decision through DetectorSession.predict_one (10:113-199) ----------------
        rngz = np.random.default_rng(5)
        names = [p[0] for p in plan if p[4] == 3.0 and p[5] == 144000]
        Zs = {k: np.load(GOLDEN / f"feat_{k}.npz")["z"] for k in names}
        Zmat = np.stack([Zs[k] for k in names])
        # centroid k sits near case k (species 3 near case 1 as well -> a multi-accept / priority case);
        # thresholds 1.3 x the nearest distance -> the remaining cases are NO_DETECT
        near = [0, 1, 2, 1]
        cent = {sp: (Zmat[near[i]] + 0.5 * rngz.standard_normal(Zmat.shape[1])).astype(np.float32)
                for i, sp in enumerate(SPECIES)}
        d_all = np.array([[np.sqrt(np.sum((Zs[k] - cent[sp]) ** 2)) for sp in SPECIES] for k in names])
        thr = {sp: float(1.3 * np.min(d_all[:, i])) for i, sp in enumerate(SPECIES)}
        cfg = {"species": SPECIES, "chunk_seconds": 3.0,
               "radial_detector": {"centroids": {sp: cent[sp].tolist() for sp in SPECIES}, "thresholds": thr}}
        cfg_path = td / "config.json"
        cfg_path.write_text(json.dumps(cfg))
        c2, t2, dur = ref09.get_detector_from_config(ref09.load_json(cfg_path))
        sess = ref10.DetectorSession(module=ref09, project_root=td, config_path=cfg_path, encoder_pt=td / "x.pt",
                                     encoder_yaml=td / "x.yaml", device="cpu")
        sess.centroids, sess.thresholds, sess.duration, sess.encoder = c2, t2, dur, encoder
        dec = {}
        for k in names:
            det, sp, best = sess.predict_one(td / f"{k}.wav")
            dec[k] = dict(detected=bool(det), species=sp, best_d=float(best))
        np.savez_compressed(GOLDEN / "decision.npz", centroids=np.stack([cent[sp] for sp in SPECIES]),
                            thresholds=np.array([thr[sp] for sp in SPECIES], dtype=np.float64))
        meta["decision"] = dec
        meta["decision_cases"] = names
    meta["feat_cases"] = feat_cases

    # ---------------- D. fit_species_with_fp_control (08:310-333) ----------------
    fits = {}
    for tag, n, d, seed in (("small", 400, 128, 31), ("tiny", 7, 16, 32), ("big", 20000, 128, 33)):
        rng = np.random.default_rng(seed)
        cents = 3.0 * rng.standard_normal((4, d))
        labels = np.arange(n) % 4
        Z = (cents[labels] + rng.standard_normal((n, d))).astype(np.float32)
        for q_out in (0.01, 0.10, 0.15, 0.20, 0.25):
            for k in range(4):
                mu, rk, rk_in, rk_out, extra = ref08.fit_species_with_fp_control(Z[labels == k], Z[labels != k], 0.95, q_out)
                fits[f"{tag}_q{q_out:.2f}_k{k}"] = dict(mu_sha=sha(mu), mu_head=mu[:4].tolist(), rk=rk, rk_in=rk_in,
                                                        rk_out=rk_out, extra=extra)
        mu, rk, rk_in, rk_out, extra = ref08.fit_species_with_fp_control(Z[labels == 0], None, 0.95, 0.1)
        fits[f"{tag}_noout"] = dict(mu_sha=sha(mu), mu_head=mu[:4].tolist(), rk=rk, rk_in=rk_in,
                                    rk_out=None if not np.isfinite(rk_out) else rk_out, extra=extra)
    meta["fit"] = fits
    meta["fit_recipe"] = "rng=default_rng(seed); cents=3*N(0,1)[4,d]; labels=arange(n)%4; Z=(cents[labels]+N(0,1)).astype(f32)"

    (GOLDEN / "meta.json").write_text(json.dumps(meta, indent=1, default=float))
    total = sum(p.stat().st_size for p in GOLDEN.iterdir())
    print(f"wrote {len(list(GOLDEN.iterdir()))} files, {total / 1e6:.2f} MB, into {GOLDEN}")


if __name__ == "__main__":
    main()
