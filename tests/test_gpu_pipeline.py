"""GPU: the file-level pipeline against the artefacts the reference's own scripts
produced on the same seeded WAV tree (tests/golden/pipeline/, oracle/make_golden_pipeline.py):
00 process_folder -> 08 fit (+cache, max_per_class sampling) -> 10 benchmark, over the q_out grid of run_qout_grid.sh."""
import csv
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = Path(__file__).parent / "golden" / "pipeline"
TOL = 1e-3            # BASELINE.json north_star: latents / radii within 1e-3, decisions equal away from thresholds


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    from amphibian_vae_latent_detector_b200 import reference_api as api
    from amphibian_vae_latent_detector_b200 import synth
    meta = json.loads((GOLD / "meta.json").read_text())
    root = tmp_path_factory.mktemp("proj")
    lse = root / "latent_space_exploration"
    synth.write_wav_tree(lse / "raw" / "train_chunks", meta["species"], meta["n_train"], meta["length"],
                         seed=meta["seed_train"], special_every=17)
    synth.write_wav_tree(lse / "raw" / "val_chunks", meta["species"], meta["n_val"], meta["length"],
                         seed=meta["seed_val"], special_every=11)
    for split in ("train_chunks", "val_chunks"):
        api.process_folder(lse / "raw" / split, lse / split, sr=48000)
    (root / "config.json").write_text(json.dumps({"species": meta["species"], "chunk_seconds": 3.0}, indent=2))
    return root, lse, meta


def test_process_folder_files_are_bit_identical(tree):
    """R2 (00:41-57): every normalised PCM_16 WAV equals the file the reference wrote, byte for byte."""
    root, lse, meta = tree
    assert len(meta["normalised_wav_sha256"]) == 4 * (meta["n_train"] + meta["n_val"])
    for rel, sha in meta["normalised_wav_sha256"].items():
        assert hashlib.sha256((lse / rel).read_bytes()).hexdigest() == sha, rel


def _read_csv(path):
    with open(path, newline="", encoding="utf-8") as f:
        return list(csv.DictReader(f))


def test_qout_grid_matches_reference_artifacts(tree):
    from amphibian_vae_latent_detector_b200 import pipeline
    from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder
    root, lse, meta = tree
    enc = build_standin_encoder(seed=123)
    grid = [float(q) for q in meta["grid"]]
    out = pipeline.run_qout_grid(lse / "train_chunks", lse / "val_chunks", root / "config.json", enc, root / "grid",
                                 q_in=float(meta["q_in"]), grid=grid, max_per_class=int(meta["max_per_class"]), seed=123)
    assert sorted(out) == sorted(meta["grid"])
    for q in meta["grid"]:
        got_dir, ref_dir = root / "grid" / f"qout_{q}", GOLD / f"qout_{q}"
        for name in ("run.log", "summary.txt", "results.csv", "config_used.json", "config_snapshot.json"):
            assert (got_dir / name).exists(), name
        got = json.loads((got_dir / "config_used.json").read_text())["radial_detector"]
        ref = json.loads((ref_dir / "config_used.json").read_text())["radial_detector"]
        names = list(ref["centroids"])
        assert list(got["centroids"]) == names                       # fit order = JSON order = decision order (09:416)
        thr_ref = np.array([ref["thresholds"][sp] for sp in names])
        thr_got = np.array([got["thresholds"][sp] for sp in names])
        assert np.max(np.abs(thr_got - thr_ref) / thr_ref) <= TOL
        for sp in names:
            a, b = np.array(got["centroids"][sp]), np.array(ref["centroids"][sp])
            assert np.max(np.abs(a - b)) / np.max(np.abs(b)) <= TOL     # same random.sample subset, same latents
            pg, pr = got["meta_fit"]["per_species"][sp], ref["meta_fit"]["per_species"][sp]
            assert list(pg) == list(pr)
            for k in ("N_in", "N_out", "failed", "used"):
                assert pg[k] == pr[k]
            for k in ("rk_in", "rk_out", "rk_final"):
                assert abs(pg[k] - pr[k]) <= TOL * abs(pr[k])
            for side in ("rho_in_summary", "rho_out_summary"):
                for k in ("min", "p50", "p90", "max"):
                    assert abs(pg[side][k] - pr[side][k]) <= TOL * max(abs(pr[side][k]), 1.0)
        assert {k: v for k, v in got["meta_fit"].items() if k not in ("per_species", "chunks_dir")} == \
               {k: v for k, v in ref["meta_fit"].items() if k not in ("per_species", "chunks_dir")}
        # decisions: identical except files whose best distance sits within TOL of a threshold
        rg, rr = _read_csv(got_dir / "results.csv"), _read_csv(ref_dir / "results.csv")
        assert [Path(r["file"]).name for r in rg] == [Path(r["file"]).name for r in rr]
        flips = 0
        for a, b in zip(rg, rr):
            da, db = float(a["best_distance"]), float(b["best_distance"])
            assert abs(da - db) <= TOL * db
            near = np.any(np.abs(db - thr_ref) / thr_ref <= 2 * TOL)
            if not near:
                assert (a["pred_species"], a["detected"], a["correct"]) == (b["pred_species"], b["detected"], b["correct"])
            else:
                flips += a["pred_species"] != b["pred_species"]
        if flips == 0:
            assert (got_dir / "summary.txt").read_text(encoding="utf-8") == (ref_dir / "summary.txt").read_text(encoding="utf-8")
        snap_g = json.loads((got_dir / "config_snapshot.json").read_text())
        snap_r = json.loads((ref_dir / "config_snapshot.json").read_text())
        assert set(snap_g) - {"timestamp"} == set(snap_r)
        for sp in names:
            assert abs(snap_g["rk_per_species"][sp] - snap_r["rk_per_species"][sp]) <= TOL * snap_r["rk_per_species"][sp]


def test_fit_radial_detector_cache_roundtrip(tree):
    """08 --cache: cache_npz/Z_<root>_<species>.npz with Z / failed / root (08:467-475, :518-520); a second run loads it."""
    from amphibian_vae_latent_detector_b200 import pipeline
    from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder
    root, lse, meta = tree
    cfgp = root / "config_fit.json"
    cfgp.write_text(json.dumps({"species": meta["species"], "chunk_seconds": 3.0}))
    enc = build_standin_encoder(seed=123)
    logs1, logs2 = [], []
    c1 = pipeline.fit_radial_detector(cfgp, lse / "train_chunks", enc, q_in=0.95, q_out=0.10, max_per_class=8, seed=123,
                                      cache=True, cache_dir=root / "cache_npz", log=logs1.append)
    assert sorted(p.name for p in (root / "cache_npz").glob("*.npz")) == meta["cache_files"]
    z = np.load(root / "cache_npz" / meta["cache_files"][0])
    assert sorted(z.files) == meta["cache_keys"] and list(z["Z"].shape) == meta["cache_Z_shape"]
    c2 = pipeline.fit_radial_detector(cfgp, lse / "train_chunks", enc, q_in=0.95, q_out=0.10, max_per_class=8, seed=123,
                                      cache=True, cache_dir=root / "cache_npz", log=logs2.append)
    assert any(ln.startswith("🧊") for ln in logs2) and not any(ln.startswith("🧊") for ln in logs1)
    assert c1["radial_detector"]["thresholds"] == c2["radial_detector"]["thresholds"]
    assert cfgp.with_suffix(".json.bak").exists()
    ref = json.loads((GOLD / "qout_0.10" / "config_used.json").read_text())["radial_detector"]["thresholds"]
    for sp, v in ref.items():
        assert abs(c1["radial_detector"]["thresholds"][sp] - v) <= TOL * v
