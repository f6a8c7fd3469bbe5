"""GPU: long-recording windows (row N4).  Window i of the stream must give exactly what the chunk-file workflow gives
for the same samples (00 normalise + PCM_16 write, then 09/10 decide), and agree with the CPU oracle."""
import json
import wave

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SPECIES = ["Batrachyla_leptopus", "Batrachyla_taeniata", "Calyptocephalella_gayi", "Pleurodema_thaul"]
L = 144000


def _write(path, pcm):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(48000)
        w.writeframes(pcm.astype("<i2").tobytes())


@pytest.fixture(scope="module")
def recording(tmp_path_factory):
    from amphibian_vae_latent_detector_b200 import synth
    from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder
    from oracle import hotpath as hp
    root = tmp_path_factory.mktemp("stream")
    n_full, tail = 9, 50000
    x, label = synth.make_chunks(n_full + 1, L, seed=77, special_every=5)
    pcm = torch.clamp(torch.round(x * 32767.0), -32768, 32767).to(torch.int16).numpy()
    long = np.concatenate([pcm[:n_full].reshape(-1), pcm[n_full, :tail]])
    _write(root / "long.wav", long)
    enc = build_standin_encoder(seed=123)
    # oracle: every window as its own chunk (the last one right-padded with zeros)
    win = np.zeros((n_full + 1, L), np.float32)
    win[:n_full] = pcm[:n_full].astype(np.float32) / 32768.0
    win[n_full, :tail] = pcm[n_full, :tail].astype(np.float32) / 32768.0
    mel_kw = dict(sr=48000, n_mels=64, fmin=150.0, fmax=15000.0, hop_length=384, n_fft=2048, target_frames=192)
    yo, oko, _ = hp.rms_normalize_batch(win, pcm16=True)
    Zo = hp.encode_batch(enc, yo, **mel_kw)
    lab = np.arange(n_full + 1) % 4
    cent, rk, _, _ = hp.fit_radial(Zo, lab, 4, 0.95, 0.25)
    rk = rk * 1.15                                   # a few accepts, a few NO_DETECT
    cfg = {"species": SPECIES, "chunk_seconds": 3.0,
           "radial_detector": {"centroids": {sp: cent[i].tolist() for i, sp in enumerate(SPECIES)},
                               "thresholds": {sp: float(rk[i]) for i, sp in enumerate(SPECIES)}}}
    (root / "config.json").write_text(json.dumps(cfg))
    pred_o, best_o, radii_o = hp.decide_batch(Zo, SPECIES, cent, rk)
    return dict(root=root, enc=enc, pcm=pcm, long=long, n_full=n_full, tail=tail, oko=oko, pred_o=pred_o, best_o=best_o,
                radii_o=radii_o, rk=rk, win=win)


def test_windows_match_oracle(recording):
    from amphibian_vae_latent_detector_b200 import stream
    r = recording
    res = stream.detect_long_wav(r["root"] / "long.wav", config_path=r["root"] / "config.json", encoder=r["enc"])
    assert len(res) == r["n_full"] + 1
    assert [w.start_s for w in res] == [3.0 * i for i in range(r["n_full"] + 1)]
    assert [w.normalised for w in res] == [bool(v) for v in r["oko"]]
    for i, w in enumerate(res):
        assert abs(w.best_distance - r["best_o"][i]) <= 1e-3 * r["best_o"][i]
        near = np.any(np.abs(r["radii_o"][i] - r["rk"]) / r["rk"] <= 1e-3)
        if not near:
            want = SPECIES[r["pred_o"][i]] if r["pred_o"][i] >= 0 else None
            assert w.species == want and w.detected == (want is not None)


def test_window_equals_chunk_file_workflow(recording, tmp_path):
    """stream window i == process_folder (00) on a chunk file with the same samples, then DetectorSession (10)."""
    from amphibian_vae_latent_detector_b200 import reference_api as api
    from amphibian_vae_latent_detector_b200 import stream
    r = recording
    raw, norm = tmp_path / "raw" / "sp", tmp_path / "norm"
    raw.mkdir(parents=True)
    for i in range(r["n_full"]):
        _write(raw / f"w{i:02d}.wav", r["pcm"][i])
    pad = np.zeros(L, np.int16)
    pad[:r["tail"]] = r["pcm"][r["n_full"], :r["tail"]]
    _write(raw / f"w{r['n_full']:02d}.wav", pad)
    api.process_folder(tmp_path / "raw", norm)
    sess = api.DetectorSession(None, r["root"], r["root"] / "config.json", device="cuda")
    sess.centroids, sess.thresholds, sess.duration = api.get_detector_from_config(api.load_json(r["root"] / "config.json"))
    sess.encoder = r["enc"]
    files = sorted((norm / "sp").glob("*.wav"))
    by_file = sess.predict_many(files)
    res = stream.detect_long_wav(r["root"] / "long.wav", config_path=r["root"] / "config.json", encoder=r["enc"])
    for w, (det, sp, best) in zip(res, by_file):
        assert (w.detected, w.species) == (det, sp)
        assert abs(w.best_distance - best) <= 1e-5 * best     # same kernels, different batch composition only


def test_overlapping_hop_and_slabs(recording):
    from amphibian_vae_latent_detector_b200 import reference_api as api
    from amphibian_vae_latent_detector_b200 import stream
    r = recording
    cents, thr, _ = api.get_detector_from_config(api.load_json(r["root"] / "config.json"))
    a = stream.detect_pcm16_stream(r["long"], r["enc"], cents, thr, window_seconds=3.0, hop_seconds=1.5, slab_windows=4)
    b = stream.detect_pcm16_stream(r["long"], r["enc"], cents, thr, window_seconds=3.0, hop_seconds=1.5, slab_windows=1000)
    assert len(a) == len(b) == len(stream.window_starts(r["long"].shape[0], L, L // 2))
    assert [(w.start_s, w.detected, w.species, w.best_distance) for w in a] == \
           [(w.start_s, w.detected, w.species, w.best_distance) for w in b]
    whole = stream.detect_pcm16_stream(r["long"], r["enc"], cents, thr, window_seconds=3.0)
    assert [(w.detected, w.species, w.best_distance) for w in a[::2]] == [(w.detected, w.species, w.best_distance) for w in whole]
