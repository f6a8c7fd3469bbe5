"""TEST INFRASTRUCTURE ONLY -- MAP-detector fixtures from the reference's own code.

    python -m oracle.make_golden_map        (build container only; needs /root/reference)

Runs the unmodified reference functions ``estimate_cov`` (08b_fit_map_detector.py:60-81), ``inv_and_logdet`` and
``gaussian_logpdf_from_precision`` (map_detector_core.py:306-323) and the decision loop of
``MapDetectorSession.predict_one`` (10b_benchmark_folder_detection_map.py:146-169, executed verbatim on latents by
substituting the encode step) on seeded synthetic latents, and stores inputs recipe + outputs in tests/golden/map.npz."""
from __future__ import annotations

import importlib.util
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))
from oracle import ref_import, shims  # noqa: E402

GOLDEN = REPO / "tests" / "golden"


def latents(n=600, d=32, seed=41, k=4):
    rng = np.random.default_rng(seed)
    cents = 2.0 * rng.standard_normal((k, d))
    A = rng.standard_normal((d, d)) * 0.3 + np.eye(d)
    labels = rng.integers(0, k, n)
    Z = (cents[labels] + rng.standard_normal((n, d)) @ A.T).astype(np.float32)
    return Z, labels.astype(np.int32)


def main():
    if not ref_import.available():
        raise SystemExit("reference tree not available")
    shims.install()
    core = ref_import.load("core")
    sys.path.insert(0, str(ref_import.REFERENCE_ROOT))
    spec = importlib.util.spec_from_file_location("_ref_08b", str(ref_import.LSE / "08b_fit_map_detector.py"))
    m08b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m08b)

    species = ["Batrachyla_leptopus", "Batrachyla_taeniata", "Calyptocephalella_gayi", "Pleurodema_thaul"]
    out = {}
    meta = {"recipe": "oracle.make_golden_map.latents(n, d, seed): cents=2*N(0,1)[k,d]; A=0.3*N(0,1)[d,d]+I; "
                      "labels=rng.integers(0,k,n); Z=(cents[labels]+N(0,1)[n,d]@A.T).astype(f32)", "cases": {}}
    for tag, n, d, seed, cov_type, structure, shrink, eps in (("lda_full", 600, 32, 41, "lda", "full", 0.0, 1e-6),
                                                               ("qda_full", 900, 16, 42, "qda", "full", 0.1, 1e-6),
                                                               ("lda_diag", 500, 128, 43, "lda", "diag", 0.0, 1e-3),
                                                               ("qda_d128", 4000, 128, 44, "qda", "full", 0.05, 1e-6)):
        Z, lab = latents(n, d, seed)
        Zs = {sp: Z[lab == i] for i, sp in enumerate(species)}
        tot = float(n)
        priors = {sp: Zs[sp].shape[0] / tot for sp in species}
        means = {sp: np.mean(Zs[sp], axis=0).astype(np.float32) for sp in species}
        covs, precs, lds = {}, {}, {}
        if cov_type == "lda":
            Zc = np.concatenate([Zs[sp] - means[sp][None] for sp in species])
            cov = m08b.estimate_cov(Zc, eps=eps, shrink=shrink, cov_structure=structure)
            prec, ld = core.inv_and_logdet(cov)
            for sp in species:
                covs[sp], precs[sp], lds[sp] = cov, prec, ld
        else:
            for sp in species:
                covs[sp] = m08b.estimate_cov(Zs[sp] - means[sp][None], eps=eps, shrink=shrink, cov_structure=structure)
                precs[sp], lds[sp] = core.inv_and_logdet(covs[sp])
        scores = np.zeros((n, 4))
        for r in range(n):
            for i, sp in enumerate(species):
                scores[r, i] = core.gaussian_logpdf_from_precision(Z[r], means[sp], precs[sp], lds[sp]) + \
                    float(np.log(priors[sp] + 1e-12))
        true_scores = scores[np.arange(n), lab]
        tau = float(np.quantile(true_scores, 0.05))
        pred = np.where(scores.max(axis=1) >= tau, scores.argmax(axis=1), -1)   # species sorted == index order here
        out[f"{tag}_cov"] = np.stack([covs[sp] for sp in species])
        out[f"{tag}_prec"] = np.stack([precs[sp] for sp in species])
        out[f"{tag}_logdet"] = np.array([lds[sp] for sp in species])
        out[f"{tag}_means"] = np.stack([means[sp] for sp in species])
        out[f"{tag}_scores"] = scores
        out[f"{tag}_pred"] = pred.astype(np.int32)
        meta["cases"][tag] = dict(n=n, d=d, seed=seed, cov_type=cov_type, cov_structure=structure, shrink=shrink, eps=eps,
                                  tau=tau, tau_q=0.05)
    np.savez_compressed(GOLDEN / "map.npz", **out)
    (GOLDEN / "map_meta.json").write_text(json.dumps(meta, indent=1))
    print("wrote", GOLDEN / "map.npz", (GOLDEN / "map.npz").stat().st_size)


if __name__ == "__main__":
    main()
