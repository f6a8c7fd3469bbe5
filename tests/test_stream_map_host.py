"""CPU: long-recording windows decided 09n / 10b style (BASELINE configs[4]) -- the host logic of
``stream.detect_long_wav_map`` (window cutting, right zero-padding, slabs, config parsing, class selection, tau) with the two
device calls it makes (``Engine.encode_detect_host`` for the latent means, ``Engine.map_score``) replaced by the numpy
oracle.  Expected values: the per-window loop of 09n:114-140 (``oracle.hotpath.decide_map_one``) on the oracle's latents."""
import json
import wave

import numpy as np
import pytest
import torch

from amphibian_vae_latent_detector_b200 import reference_api as api
from amphibian_vae_latent_detector_b200 import stream, synth
from oracle import hotpath as hp
from test_map_cli_host import OracleEngine

SPECIES = ["Batrachyla_leptopus", "Batrachyla_taeniata", "Calyptocephalella_gayi", "Pleurodema_thaul"]
L = 144000
MEL = dict(sr=48000, n_mels=64, fmin=150.0, fmax=15000.0, hop_length=384, n_fft=2048, target_frames=192)


class OracleStreamEngine(OracleEngine):
    latent_dim = 128

    def __init__(self, encoder):
        self.encoder = encoder
        self.calls = []

    def encode_detect_host(self, x_host, centroid, thr, priority_rank, *, pcm16=True, want_mu=False):
        x = x_host.numpy() if isinstance(x_host, torch.Tensor) else np.asarray(x_host)
        assert centroid.shape == (1, 128) and np.isinf(thr).all() and want_mu and pcm16
        self.calls.append(x.shape[0])
        if x.dtype == np.int16:
            x = x.astype(np.float32) / np.float32(32768.0)
        y, ok, _ = hp.rms_normalize_batch(x, pcm16=True)
        mu = hp.encode_batch(self.encoder, y, **MEL)
        return np.zeros(x.shape[0], np.int32), np.linalg.norm(mu, axis=1).astype(np.float32), ok.astype(np.uint8), mu


@pytest.fixture(scope="module")
def recording(tmp_path_factory, standin_encoder):
    root = tmp_path_factory.mktemp("stream_map")
    n_full, tail = 5, 50000
    x, _ = synth.make_chunks(n_full + 1, L, seed=77, special_every=5)
    pcm = torch.clamp(torch.round(x * 32767.0), -32768, 32767).to(torch.int16).numpy()
    long = np.concatenate([pcm[:n_full].reshape(-1), pcm[n_full, :tail]])
    with wave.open(str(root / "long.wav"), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(48000)
        w.writeframes(long.astype("<i2").tobytes())
    win = np.zeros((n_full + 1, L), np.float32)
    win[:n_full] = pcm[:n_full].astype(np.float32) / 32768.0
    win[n_full, :tail] = pcm[n_full, :tail].astype(np.float32) / 32768.0
    yo, oko, _ = hp.rms_normalize_batch(win, pcm16=True)
    Zo = hp.encode_batch(standin_encoder, yo, **MEL)
    rng = np.random.default_rng(0)
    Ztrain = {sp: Zo[i] + 0.3 * rng.standard_normal((40, 128)).astype(np.float32) for i, sp in enumerate(SPECIES)}
    fit = hp.fit_map(Ztrain, cov_type="lda", eps=1e-3, set_tau_q=0.05)
    cfg = {"species": SPECIES, "chunk_seconds": 5.0, "map_detector": {
        "model": "gaussian_map", "means": {sp: fit["means"][sp].tolist() for sp in SPECIES},
        "precision": {sp: fit["precision"][sp].tolist() for sp in SPECIES},
        "logdet_cov": {sp: fit["logdet_cov"][sp] for sp in SPECIES}, "tau": fit["tau"],
        "meta_fit": {"chunk_seconds": 3.0, "per_species": {sp: {"prior": fit["priors"][sp]} for sp in SPECIES}}}}
    (root / "config.json").write_text(json.dumps(cfg))
    want = [hp.decide_map_one(z, fit["species"], fit["means"], fit["precision"], fit["logdet_cov"], fit["priors"], fit["tau"])
            for z in Zo]
    return dict(root=root, long=long, n=n_full + 1, oko=oko, want=want, fit=fit)


@pytest.fixture()
def oracle_engine(monkeypatch, standin_encoder):
    eng = OracleStreamEngine(standin_encoder)
    monkeypatch.setattr(api, "_engine_with_encoder", lambda *a, **k: eng)
    return eng


@pytest.mark.parametrize("slab", [4096, 2])
def test_windows_follow_09n(recording, oracle_engine, standin_encoder, slab):
    r = recording
    res = stream.detect_long_wav_map(r["root"] / "long.wav", config_path=r["root"] / "config.json", encoder=standin_encoder,
                                     slab_windows=slab)
    assert [w.start_s for w in res] == [3.0 * i for i in range(r["n"])]            # meta_fit.chunk_seconds wins (core:358-370)
    assert [w.normalised for w in res] == [bool(v) for v in r["oko"]]
    assert oracle_engine.calls == ([r["n"]] if slab >= r["n"] else [2, 2, 2])
    kinds = set()
    for w, (det, sp, best) in zip(res, r["want"]):
        assert w.best_score == pytest.approx(best, rel=1e-5, abs=1e-3)
        assert (w.detected, w.species) == (det, sp)
        kinds.add(w.detected)
    assert kinds == {True, False}                                                  # accepted windows and rejected ones


def test_overlap_and_empty_class_set(recording, oracle_engine, standin_encoder):
    r = recording
    cfg = api.load_json(r["root"] / "config.json")
    means, precs, lds, tau = api.read_map_detector_params(cfg)
    pri = api.get_priors_from_map_meta(cfg, sorted(means))
    whole = stream.detect_pcm16_stream_map(r["long"], standin_encoder, means, precs, lds, pri, tau, window_seconds=3.0)
    half = stream.detect_pcm16_stream_map(r["long"], standin_encoder, means, precs, lds, pri, tau, window_seconds=3.0,
                                          hop_seconds=1.5, slab_windows=3)
    assert len(half) == len(stream.window_starts(r["long"].shape[0], L, L // 2))
    assert [(w.detected, w.species, w.best_score) for w in half[::2]] == [(w.detected, w.species, w.best_score) for w in whole]
    # latent size of the config does not match the encoder: every class is skipped (09n:120-123) -> NO_DETECT, -inf (09n:142-143)
    bad = {sp: m[:64] for sp, m in means.items()}
    none = stream.detect_pcm16_stream_map(r["long"], standin_encoder, bad, precs, lds, pri, tau, window_seconds=3.0)
    assert len(none) == r["n"] and all((not w.detected) and w.species is None and w.best_score == -float("inf") for w in none)
    with pytest.raises(ValueError):
        stream.detect_pcm16_stream_map(r["long"], standin_encoder, means, precs, lds, pri, tau, hop_seconds=0.0)


def test_other_sample_formats_take_the_float_path(recording, oracle_engine, standin_encoder, tmp_path):
    """A stereo file cannot be mapped as mono PCM_16: decoded as librosa.load would (channel mean), then the same windows."""
    r = recording
    st = np.stack([r["long"], r["long"]], axis=1).astype("<i2")
    with wave.open(str(tmp_path / "stereo.wav"), "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(48000)
        w.writeframes(st.tobytes())
    res = stream.detect_long_wav_map(tmp_path / "stereo.wav", config_path=r["root"] / "config.json", encoder=standin_encoder)
    for w, (det, sp, best) in zip(res, r["want"]):
        assert (w.detected, w.species) == (det, sp) and w.best_score == pytest.approx(best, rel=1e-5, abs=1e-3)
