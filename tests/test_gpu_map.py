"""GPU, row N1: Gaussian-MAP fit + scoring + decision vs fixtures made by the reference's own functions
(tests/golden/map.npz) and vs the numpy oracle."""
import json

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import hotpath as hp
from oracle.make_golden_map import latents

pytestmark = pytest.mark.gpu
SPECIES = ["Batrachyla_leptopus", "Batrachyla_taeniata", "Calyptocephalella_gayi", "Pleurodema_thaul"]


@pytest.mark.parametrize("tag", ["lda_full", "qda_full", "lda_diag", "qda_d128"])
def test_map_fit_scores_decisions_vs_reference(engine3s, tag):
    g = np.load(GOLDEN / "map.npz")
    c = json.loads((GOLDEN / "map_meta.json").read_text())["cases"][tag]
    Z, lab = latents(c["n"], c["d"], c["seed"])
    Zd, ld = torch.from_numpy(Z).cuda(), torch.from_numpy(lab).cuda()
    fit = engine3s.fit_map(Zd, ld, SPECIES, cov_type=c["cov_type"], cov_structure=c["cov_structure"], eps=c["eps"],
                           shrink=c["shrink"], set_tau_q=c["tau_q"])
    assert fit.species == SPECIES
    # np.mean(axis=0) of float32 rows adds sequentially in float32; the GPU sum is float64
    assert np.allclose(fit.means, g[f"{tag}_means"], rtol=2e-5, atol=1e-5)
    scale = np.abs(g[f"{tag}_cov"]).max()
    assert np.max(np.abs(fit.cov - g[f"{tag}_cov"])) <= 1e-5 * scale
    assert np.allclose(fit.logdet_cov, g[f"{tag}_logdet"], rtol=1e-4, atol=1e-3)
    # scores: quadratic form in float32 on both sides, different summation order
    pred, best, scores = engine3s.map_score(Zd, fit, want_scores=True)
    ref = g[f"{tag}_scores"]
    assert np.max(np.abs(scores.cpu().numpy() - ref)) <= 1e-3 * np.max(np.abs(ref))
    assert fit.tau == pytest.approx(c["tau"], rel=1e-3)
    # decisions identical except latents within tolerance of a tie / of tau
    top2 = np.sort(ref, axis=1)[:, -2:]
    near = (np.abs(top2[:, 1] - top2[:, 0]) < 1e-2) | (np.abs(top2[:, 1] - c["tau"]) < 1e-2 * abs(c["tau"]))
    p = pred.cpu().numpy()
    assert np.array_equal(p[~near], g[f"{tag}_pred"][~near])
    assert near.mean() < 0.2 and (p == -1).sum() > 0
    assert np.allclose(best.cpu().numpy(), ref.max(axis=1), rtol=1e-3)


def test_map_score_with_given_precision_is_tight(engine3s):
    """Same parameters on both sides (the reference-made precision matrices): only the kernel differs."""
    from amphibian_vae_latent_detector_b200.map_fit import MapFit
    g = np.load(GOLDEN / "map.npz")
    c = json.loads((GOLDEN / "map_meta.json").read_text())["cases"]["qda_d128"]
    Z, lab = latents(c["n"], c["d"], c["seed"])
    counts = np.bincount(lab, minlength=4)
    fit = MapFit(SPECIES, np.arange(4, dtype=np.int32), g["qda_d128_means"], g["qda_d128_cov"], g["qda_d128_prec"],
                 g["qda_d128_logdet"], counts / counts.sum(), c["tau"], counts)
    pred, best, scores = engine3s.map_score(torch.from_numpy(Z).cuda(), fit, want_scores=True)
    ref = g["qda_d128_scores"]
    assert np.max(np.abs(scores.cpu().numpy() - ref) / np.abs(ref)) < 2e-5
    top2 = np.sort(ref, axis=1)[:, -2:]
    near = (np.abs(top2[:, 1] - top2[:, 0]) < 1e-2) | (np.abs(top2[:, 1] - c["tau"]) < 1e-2)
    assert np.array_equal(pred.cpu().numpy()[~near], g["qda_d128_pred"][~near])
    # the per-latent oracle path (09n decision loop) on a few rows
    means = {sp: fit.means[i] for i, sp in enumerate(SPECIES)}
    precs = {sp: fit.precision[i] for i, sp in enumerate(SPECIES)}
    lds = {sp: float(fit.logdet_cov[i]) for i, sp in enumerate(SPECIES)}
    pri = {sp: float(fit.priors[i]) for i, sp in enumerate(SPECIES)}
    for r in (0, 5, 77, 1234):
        det, sp, b = hp.decide_map_one(Z[r], SPECIES, means, precs, lds, pri, c["tau"])
        assert float(best[r]) == pytest.approx(b, rel=2e-5)


def test_reference_named_map_api(tmp_path, standin_encoder):
    """MapDetectorSession / estimate_cov / gaussian_logpdf_from_precision with the reference's names (10b, 08b, core)."""
    from amphibian_vae_latent_detector_b200 import reference_api as api
    from conftest import load_pcm_case
    from oracle import librosa_port as lp
    Z, lab = latents(600, 32, 41)
    assert np.max(np.abs(api.estimate_cov(Z[lab == 1], 1e-6, 0.1, "full") - hp.estimate_cov(Z[lab == 1], 1e-6, 0.1, "full"))) < 1e-5
    g = np.load(GOLDEN / "map.npz")
    s = api.gaussian_logpdf_from_precision(Z[3], g["lda_full_means"][0], g["lda_full_prec"][0], float(g["lda_full_logdet"][0]))
    assert s == pytest.approx(hp.gaussian_logpdf_from_precision(Z[3], g["lda_full_means"][0], g["lda_full_prec"][0],
                                                                 float(g["lda_full_logdet"][0])), rel=2e-5)
    # a MAP config fitted on latents of golden chunks, then evaluated on the WAV files through the session
    keys = ["noise_3s", "tonal_3s", "pulsed_3s", "burst0_3s", "burst3_3s", "hot_3s"]
    zs = np.stack([np.load(GOLDEN / f"feat_{k}.npz")["z"] for k in keys])
    rng = np.random.default_rng(0)
    Ztrain = np.concatenate([zs[i] + 0.3 * rng.standard_normal((40, 128)).astype(np.float32) for i in range(4)])
    ltrain = np.repeat(np.arange(4), 40)
    fit = hp.fit_map({sp: Ztrain[ltrain == i] for i, sp in enumerate(SPECIES)}, cov_type="lda", eps=1e-3, set_tau_q=0.05)
    cfg = {"chunk_seconds": 3.0, "map_detector": {
        "model": "gaussian_map", "means": {sp: fit["means"][sp].tolist() for sp in SPECIES},
        "precision": {sp: fit["precision"][sp].tolist() for sp in SPECIES},
        "logdet_cov": {sp: fit["logdet_cov"][sp] for sp in SPECIES}, "tau": fit["tau"],
        "meta_fit": {"per_species": {sp: {"prior": fit["priors"][sp]} for sp in SPECIES}}}}
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    wavs = []
    for k in keys:
        x, d = load_pcm_case(GOLDEN / f"feat_{k}.npz")
        y, _ = hp.rms_normalize(x)
        lp.write_wav(tmp_path / f"{k}.wav", np.asarray(y, np.float32), 48000)
        wavs.append(tmp_path / f"{k}.wav")
    sess = api.MapDetectorSession(tmp_path, tmp_path / "config.json", tmp_path / "x.pt", tmp_path / "x.yaml", "cuda")
    sess.set_params(cfg)
    sess.encoder = standin_encoder
    got = sess.predict_many(wavs)
    for k, z, (det, sp, best) in zip(keys, zs, got):
        d0, s0, b0 = hp.decide_map_one(z, fit["species"], fit["means"], fit["precision"], fit["logdet_cov"],
                                       fit["priors"], fit["tau"])
        assert best == pytest.approx(b0, rel=2e-2, abs=2.0)      # latents differ by <= 1e-3, scores scale with 1/var
        if abs(b0 - fit["tau"]) > 5.0:
            assert (det, sp) == (d0, s0), k
    assert got[0][:2] == (True, SPECIES[0]) and got[4][0] is False     # chunk 0 sits on species 0; chunk 4 is far from all
