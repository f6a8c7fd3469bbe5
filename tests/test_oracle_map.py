"""CPU: the oracle's MAP restatement (oracle/hotpath.py, row N1) reproduces fixtures made by the reference's own
estimate_cov / inv_and_logdet / gaussian_logpdf_from_precision (tests/golden/map.npz, oracle/make_golden_map.py)."""
import json

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import hotpath as hp
from oracle.make_golden_map import latents

SPECIES = ["Batrachyla_leptopus", "Batrachyla_taeniata", "Calyptocephalella_gayi", "Pleurodema_thaul"]


@pytest.mark.parametrize("tag", ["lda_full", "qda_full", "lda_diag", "qda_d128"])
def test_map_fit_and_scores_vs_reference(tag):
    g = np.load(GOLDEN / "map.npz")
    c = json.loads((GOLDEN / "map_meta.json").read_text())["cases"][tag]
    Z, lab = latents(c["n"], c["d"], c["seed"])
    fit = hp.fit_map({sp: Z[lab == i] for i, sp in enumerate(SPECIES)}, cov_type=c["cov_type"],
                     cov_structure=c["cov_structure"], eps=c["eps"], shrink=c["shrink"], set_tau_q=c["tau_q"])
    for i, sp in enumerate(SPECIES):
        assert np.array_equal(fit["cov"][sp], g[f"{tag}_cov"][i])
        assert np.array_equal(fit["precision"][sp], g[f"{tag}_prec"][i])
        assert fit["logdet_cov"][sp] == g[f"{tag}_logdet"][i]
        assert np.array_equal(fit["means"][sp], g[f"{tag}_means"][i])
    assert fit["tau"] == pytest.approx(c["tau"], rel=1e-12)
    for r in range(0, c["n"], 37):
        det, sp, best = hp.decide_map_one(Z[r], fit["species"], fit["means"], fit["precision"], fit["logdet_cov"],
                                          fit["priors"], fit["tau"])
        assert best == g[f"{tag}_scores"][r].max()
        assert (SPECIES.index(sp) if det else -1) == g[f"{tag}_pred"][r]
