"""CPU: the numpy oracle (oracle/hotpath.py) reproduces the outputs of the reference's own code
recorded in tests/golden/ by oracle/make_golden.py (reference run in the build container)."""
import hashlib

import numpy as np
import pytest

from conftest import GOLDEN, load_pcm_case
from oracle import hotpath as hp
from oracle import librosa_port as lp

MEL_KW = dict(sr=48000, n_mels=64, fmin=150.0, fmax=15000.0, hop_length=384, n_fft=2048, target_frames=192)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_rms_bit_exact_vs_reference(golden_meta):
    cases = golden_meta["rms_cases"]
    assert len(cases) >= 20
    n_gate = 0
    for key, info in cases.items():
        x, d = load_pcm_case(GOLDEN / f"rms_{key}.npz")
        y, ok = hp.rms_normalize(x)
        assert bool(ok) == info["ok"], key
        n_gate += (not ok)
        assert sha(np.asarray(y, dtype=np.float32)) == info["sha"], key       # bit-exact, all samples
        assert np.array_equal(np.asarray(y[:64], dtype=np.float32), d["y_head"])
        assert np.array_equal(np.asarray(y[::997], dtype=np.float32), d["y_stride"])
    assert n_gate >= 4          # the silence gate path is exercised


def test_rms_batch_matches_rowwise(golden_meta):
    rows = [load_pcm_case(GOLDEN / f"rms_{k}_144000.npz")[0] for k in ("noise", "tonal", "hot", "silent")]
    x = np.stack(rows)
    y, ok, rms = hp.rms_normalize_batch(x)
    assert ok.tolist() == [1, 1, 1, 0]
    for i, r in enumerate(rows):
        assert np.array_equal(y[i], np.asarray(hp.rms_normalize(r)[0], dtype=np.float32))
    yq, _, _ = hp.rms_normalize_batch(x, pcm16=True)
    assert np.array_equal(yq[0], lp.pcm16_roundtrip(y[0]))


@pytest.mark.parametrize("key", sorted(p.stem[5:] for p in GOLDEN.glob("feat_*.npz")))
def test_features_and_latents_vs_reference(key, standin_encoder):
    x, d = load_pcm_case(GOLDEN / f"feat_{key}.npz")
    duration = float(d["duration"])
    y, ok = hp.rms_normalize(x)
    assert int(ok) == int(d["ok"])
    y = lp.pcm16_roundtrip(np.asarray(y, dtype=np.float32))     # sf.write + librosa.load of the reference dataflow
    feat = hp.logmel_features(y, duration=duration, **MEL_KW)
    assert feat.shape == (64, 192) and feat.dtype == np.float32
    assert np.array_equal(feat, d["feat"]), key                 # same numpy ops, same machine: identical
    z = hp.encode_features(standin_encoder, feat)
    ref = d["z"]
    assert np.max(np.abs(z - ref)) <= 1e-5 * max(1.0, np.max(np.abs(ref))), key


def test_fit_vs_reference(golden_meta):
    for tag, n, dim, seed in (("small", 400, 128, 31), ("tiny", 7, 16, 32), ("big", 20000, 128, 33)):
        rng = np.random.default_rng(seed)
        cents = 3.0 * rng.standard_normal((4, dim))
        labels = np.arange(n) % 4
        Z = (cents[labels] + rng.standard_normal((n, dim))).astype(np.float32)
        for q_out in (0.01, 0.10, 0.15, 0.20, 0.25):
            for k in range(4):
                g = golden_meta["fit"][f"{tag}_q{q_out:.2f}_k{k}"]
                mu, rk, rk_in, rk_out, extra = hp.fit_species_with_fp_control(Z[labels == k], Z[labels != k], 0.95, q_out)
                assert sha(mu) == g["mu_sha"]
                assert (rk, rk_in, rk_out) == (g["rk"], g["rk_in"], g["rk_out"])
                assert extra["rho_in_summary"] == g["extra"]["rho_in_summary"]
                assert extra["rho_out_summary"] == g["extra"]["rho_out_summary"]
            # fit_radial (the 08:530-558 loop) agrees with the per-species calls
            cent, rkv, rki, rko = hp.fit_radial(Z, labels, 4, 0.95, q_out)
            for k in range(4):
                assert rkv[k] == golden_meta["fit"][f"{tag}_q{q_out:.2f}_k{k}"]["rk"]
        g = golden_meta["fit"][f"{tag}_noout"]
        mu, rk, rk_in, rk_out, _ = hp.fit_species_with_fp_control(Z[labels == 0], None, 0.95, 0.1)
        assert rk == g["rk"] and rk_out == float("inf") and g["rk_out"] is None


def test_decision_vs_reference(golden_meta):
    d = np.load(GOLDEN / "decision.npz")
    species = golden_meta["species"]
    cent = {sp: d["centroids"][i] for i, sp in enumerate(species)}
    thr = {sp: float(d["thresholds"][i]) for i, sp in enumerate(species)}
    seen = set()
    for key in golden_meta["decision_cases"]:
        z = np.load(GOLDEN / f"feat_{key}.npz")["z"]
        det, sp, best = hp.decide_one(z, cent, thr)
        g = golden_meta["decision"][key]
        assert (det, sp) == (g["detected"], g["species"]), key
        assert best == pytest.approx(g["best_d"], rel=1e-6)
        seen.add(sp)
    assert None in seen and len(seen) >= 2       # both NO_DETECT and detections are exercised
    Z = np.stack([np.load(GOLDEN / f"feat_{k}.npz")["z"] for k in golden_meta["decision_cases"]])
    pred, best, radii = hp.decide_batch(Z, species, d["centroids"], d["thresholds"])
    for i, key in enumerate(golden_meta["decision_cases"]):
        g = golden_meta["decision"][key]
        assert (species[pred[i]] if pred[i] >= 0 else None) == g["species"]


def test_priority_tie_break():
    """09:61-66, :428-436 -- several accepted species -> first in PRIORITY_ORDER; unknown names -> sorted."""
    z = np.zeros(4, dtype=np.float32)
    cent = {"Pleurodema_thaul": z.copy(), "Batrachyla_taeniata": z.copy(), "zzz": z.copy(), "aaa": z.copy()}
    thr = {k: 1.0 for k in cent}
    assert hp.decide_one(z, cent, thr) == (True, "Batrachyla_taeniata", 0.0)
    del cent["Pleurodema_thaul"], cent["Batrachyla_taeniata"]
    assert hp.decide_one(z, cent, thr) == (True, "aaa", 0.0)
    assert hp.decide_one(z + 10, cent, thr)[:2] == (False, None)
