"""GPU: the reference-named Python surface (reference_api.py) on files and arrays, against the fixtures the
reference's own functions produced (tests/golden) and the oracle."""
import hashlib
import json

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_pcm_case
from oracle import hotpath as hp
from oracle import librosa_port as lp

pytestmark = pytest.mark.gpu
MEL_KW = dict(sr=48000, n_mels=64, fmin=150.0, fmax=15000.0, hop_length=384, n_fft=2048, target_frames=192)
SPECIES = ["Batrachyla_leptopus", "Batrachyla_taeniata", "Calyptocephalella_gayi", "Pleurodema_thaul"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_rms_normalize_all_golden_lengths(golden_meta):
    """every reference-made RMS case incl. L = 7, 100, 1000, 4097 (n < 8 loop, tail loop, odd splits of the pairwise tree)."""
    from amphibian_vae_latent_detector_b200 import reference_api as api
    for key, info in golden_meta["rms_cases"].items():
        x, _ = load_pcm_case(GOLDEN / f"rms_{key}.npz")
        y, ok = api.rms_normalize(x)
        assert ok == info["ok"], key
        assert y.dtype == np.float32 and sha(y) == info["sha"], key


def test_process_folder_and_wav_to_mel_and_session(tmp_path, standin_encoder, golden_meta):
    from amphibian_vae_latent_detector_b200 import reference_api as api
    keys = ["noise_3s", "tonal_3s", "pulsed_3s", "burst0_3s", "silent_3s"]
    src = tmp_path / "train_chunks" / "Pleurodema_thaul"
    src.mkdir(parents=True)
    for k in keys:
        x, _ = load_pcm_case(GOLDEN / f"feat_{k}.npz")
        pcm = np.load(GOLDEN / f"feat_{k}.npz")
        # raw chunk files: the fixture inputs are exactly int16/32768 * gain; write those with gain 1 as PCM_16
        if float(pcm["gain"]) != 1.0:
            continue
        lp.write_wav(src / f"{k}.wav", pcm["pcm"], 48000)
    api.process_folder(tmp_path / "train_chunks", tmp_path / "train_chunks_norm", sr=48000)         # 00:41-57
    out = tmp_path / "train_chunks_norm" / "Pleurodema_thaul"
    done = sorted(p.stem for p in out.glob("*.wav"))
    assert len(done) >= 4
    for k in done:
        d = np.load(GOLDEN / f"feat_{k}.npz")
        mel = api.wav_to_mel(out / f"{k}.wav", duration=3.0, **MEL_KW)                              # core:198-237
        assert isinstance(mel, torch.Tensor) and mel.shape == (64, 192) and mel.dtype == torch.float32
        assert float(np.max(np.abs(mel.numpy() - d["feat"])) / np.max(np.abs(d["feat"]))) < 2e-4, k
        z = api.encode_wav_to_latent(standin_encoder, out / f"{k}.wav", "cuda", duration=3.0, **MEL_KW)
        assert z.shape == (128,) and z.dtype == np.float32
        assert float(np.max(np.abs(z - d["z"])) / np.max(np.abs(d["z"]))) < 1e-3, k

    # DetectorSession (10:113-199) with the decision fixture
    dec = np.load(GOLDEN / "decision.npz")
    cfg = {"species": SPECIES, "chunk_seconds": 3.0,
           "radial_detector": {"centroids": {sp: dec["centroids"][i].tolist() for i, sp in enumerate(SPECIES)},
                               "thresholds": {sp: float(dec["thresholds"][i]) for i, sp in enumerate(SPECIES)}}}
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    sess = api.DetectorSession(None, tmp_path, tmp_path / "config.json", tmp_path / "x.pt", tmp_path / "x.yaml", "cuda")
    sess.centroids, sess.thresholds, sess.duration = api.get_detector_from_config(cfg)
    sess.encoder = standin_encoder
    wavs = [out / f"{k}.wav" for k in done] + [tmp_path / "missing.wav"]
    many = sess.predict_many(wavs)
    assert many[-1][1] == "ERROR"
    for k, (det, sp, best) in zip(done, many):
        g = golden_meta["decision"][k]
        assert (det, sp) == (g["detected"], g["species"]), k
        assert best == pytest.approx(g["best_d"], rel=1e-3)
        assert sess.predict_one(out / f"{k}.wav")[:2] == (det, sp)


def test_fit_species_and_helpers_vs_oracle():
    from amphibian_vae_latent_detector_b200 import reference_api as api
    rng = np.random.default_rng(31)
    cents = 3.0 * rng.standard_normal((4, 128))
    labels = np.arange(400) % 4
    Z = (cents[labels] + rng.standard_normal((400, 128))).astype(np.float32)
    mu, rk, rk_in, rk_out, extra = api.fit_species_with_fp_control(Z[labels == 1], Z[labels != 1], 0.95, 0.10)
    mo, ro, rio, roo, eo = hp.fit_species_with_fp_control(Z[labels == 1], Z[labels != 1], 0.95, 0.10)
    assert np.allclose(mu, mo, rtol=1e-5, atol=1e-6) and mu.dtype == np.float32
    assert (rk, rk_in, rk_out) == pytest.approx((ro, rio, roo), rel=1e-5)
    for part in ("rho_in_summary", "rho_out_summary"):
        for k in ("min", "p50", "p90", "max"):
            assert extra[part][k] == pytest.approx(eo[part][k], rel=1e-5)
    _, rk2, _, rk_out2, _ = api.fit_species_with_fp_control(Z[labels == 1], None, 0.95, 0.10)
    assert rk_out2 == float("inf") and rk2 == pytest.approx(rio, rel=1e-5)
    rho = hp.l2_norm_rows(Z)
    assert np.allclose(api.l2_norm_rows(Z), rho, rtol=1e-6)
    assert api.quantile_safe(rho, 0.95) == hp.quantile_safe(rho, 0.95)        # exact selection + numpy's lerp
    assert api.quantile_safe(np.array([], dtype=np.float32), 0.5) == 0.0
    assert api.summarize_dist(rho) == hp.summarize_dist(rho)
    assert api.l2(Z[0]) == pytest.approx(hp.l2(Z[0]), rel=1e-6)


def test_auto_find_frames_and_report(tmp_path):
    """07:355-409 / :472-527 -- the frame search walks start, start + step, ... and returns the first target_frames whose
    flattened conv output matches the first Linear (stand-in encoder: 128 x (T // 16) x 4 == 6144 <=> 192 <= T < 208)."""
    import json
    import wave
    from amphibian_vae_latent_detector_b200 import reference_api as api
    from amphibian_vae_latent_detector_b200 import synth
    from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder
    x, _ = synth.make_chunks(1, 144000, seed=9, special_every=0)
    pcm = torch.clamp(torch.round(x[0] * 32767.0), -32768, 32767).to(torch.int16).numpy()
    wav = tmp_path / "a.wav"
    with wave.open(str(wav), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(48000); w.writeframes(pcm.astype("<i2").tobytes())
    enc = build_standin_encoder(seed=123)
    kw = dict(sr=48000, duration=3.0, n_mels=64, fmin=150.0, fmax=15000.0, hop_length=384, n_fft=2048)
    assert api.auto_find_frames_with_hook(enc, wav, "cpu", start_frames=192, max_frames=512, step=8, **kw) == 192
    assert api.auto_find_frames_with_hook(enc, wav, "cpu", start_frames=100, max_frames=512, step=8, **kw) == 196
    assert api.auto_find_frames_with_hook(enc, wav, "cpu", start_frames=1, max_frames=512, step=23, **kw) == 192   # 8 + 23 k
    with pytest.raises(SystemExit):                                  # 8, 58, ..., 208, 258, ...: never in [192, 208)
        api.auto_find_frames_with_hook(enc, wav, "cpu", start_frames=1, max_frames=512, step=50, **kw)
    with pytest.raises(SystemExit):
        api.auto_find_frames_with_hook(enc, wav, "cpu", start_frames=8, max_frames=150, step=8, **kw)
    lines = []
    v = api.encode_wav_report(wav, enc, device="cpu", jsonl=True, auto_frames=True, target_frames=100, log=lines.append, **kw)
    assert lines[0] == "✅ target_frames usado: 196"
    rec = json.loads(lines[1])
    assert rec["latent_dim"] == 128 and np.allclose(rec["vector"], v)
    # same vector as the batched CUDA encoder gives for the 192-frame crop?  no: 196 frames is a different crop; check 192
    v192 = api.encode_wav_report(wav, enc, device="cpu", log=lambda s: None, **kw)
    z = api.encode_wav_to_latent(enc, wav, None, duration=3.0, sr=48000, n_mels=64, fmin=150.0, fmax=15000.0,
                                 hop_length=384, n_fft=2048, target_frames=192)
    assert np.max(np.abs(v192 - z)) / np.max(np.abs(z)) < 1e-3


def test_load_wav_resamples_like_librosa_kaiser_best(tmp_path):
    """M1: a file whose rate differs from `sr` (librosa.load(path, sr=sr), core:210): avld_resample against the oracle's
    restatement of resampy kaiser_best -- the same taps in the same order, so float32 results agree to rounding of the
    filter table (1e-6 of the peak); up-sampling, down-sampling, stereo (mono mix first), odd lengths."""
    import wave
    from amphibian_vae_latent_detector_b200 import reference_api as api
    from oracle import librosa_port as lp
    rng = np.random.default_rng(5)
    for sr_in, nch, n in ((44100, 1, 30011), (96000, 1, 50000), (22050, 2, 12345), (8000, 1, 4000)):
        v = np.clip(rng.standard_normal((n, nch)) * 6000 + 8000 * np.sin(np.arange(n) * 0.05)[:, None], -32768, 32767).astype("<i2")
        path = tmp_path / f"r{sr_in}.wav"
        with wave.open(str(path), "wb") as w:
            w.setnchannels(nch); w.setsampwidth(2); w.setframerate(sr_in); w.writeframes(v.tobytes())
        got = api.load_wav(path, sr=48000)
        mono = np.mean((v.astype(np.float32) / np.float32(32768.0)).T, axis=0).astype(np.float32) if nch > 1 else \
            v[:, 0].astype(np.float32) / np.float32(32768.0)
        ref = lp.resample(mono, sr_in, 48000)
        assert got.dtype == np.float32 and got.shape == ref.shape, (got.shape, ref.shape)
        assert np.max(np.abs(got - ref)) <= 2e-6 * np.max(np.abs(ref)), (sr_in, np.max(np.abs(got - ref)))
        o, _ = lp.load(path, sr=48000)
        assert np.array_equal(o, ref)
