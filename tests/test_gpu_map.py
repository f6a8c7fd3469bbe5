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
    assert np.allclose(fit.means, g[f"{tag}_means"], rtol=1e-6, atol=1e-7)
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
