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


def test_cli_scripts_08_10_09_on_the_tree(tree, capsys, monkeypatch):
    """The reference's command lines (run_qout_grid.sh:28-38: ``08 --root train_chunks --q-in --q-out --max-per-class --seed``,
    then ``10 --root val_chunks``; 09 ``--wav`` with its 0 / 2 exit codes) through ``cli.main_*`` with the encoder loaded
    from ``downloaded_models/.../model.pt`` (a state_dict) + the YAML ``_target_``, as core:150-179 does."""
    import torch
    from amphibian_vae_latent_detector_b200 import cli
    from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder
    root, lse, meta = tree
    enc_dir = root / "downloaded_models" / "bird_net_vae_audio_splitted_encoder_v0"
    enc_dir.mkdir(parents=True, exist_ok=True)
    torch.save(build_standin_encoder(seed=123).state_dict(), enc_dir / "model.pt")
    (enc_dir / "bird_net_vae_audio_splitted.yaml").write_text(
        "encoder:\n  _target_: amphibian_vae_latent_detector_b200.encoder.build_standin_encoder\n  seed: 7\n")   # weights come from model.pt
    cfgp = root / "config_cli.json"
    cfgp.write_text(json.dumps({"species": meta["species"], "chunk_seconds": 3.0}, indent=2))
    monkeypatch.chdir(root)
    here = lse
    cli.main_08(["--config", "config_cli.json", "--root", "train_chunks", "--q-in", str(meta["q_in"]), "--q-out", "0.10",
                 "--max-per-class", str(meta["max_per_class"]), "--seed", "123", "--device", "cuda"], here=here)
    out = capsys.readouterr().out
    assert "📌 Project root:" in out and "📦 WAVs por especie en root:" in out
    got = json.loads(cfgp.read_text())["radial_detector"]
    ref = json.loads((GOLD / "qout_0.10" / "config_used.json").read_text())["radial_detector"]
    for sp, v in ref["thresholds"].items():
        assert abs(got["thresholds"][sp] - v) <= TOL * v
    # 10: folder benchmark -> <project>/outputs/detection_benchmark/{results.csv,summary.txt}
    cli.main_10(["--root", str(lse / "val_chunks"), "--config", str(cfgp)], here=here)        # default --device cpu: notice, runs on the GPU
    out = capsys.readouterr().out
    assert "solo corre en GPU" in out and "✅ CSV guardado:" in out
    rows = _read_csv(root / "outputs" / "detection_benchmark" / "results.csv")
    rr = _read_csv(GOLD / "qout_0.10" / "results.csv")
    assert [Path(r["file"]).name for r in rows] == [Path(r["file"]).name for r in rr]
    thr_ref = np.array(list(ref["thresholds"].values()))
    for a, b in zip(rows, rr):
        db = float(b["best_distance"])
        assert abs(float(a["best_distance"]) - db) <= TOL * db
        if not np.any(np.abs(db - thr_ref) / thr_ref <= 2 * TOL):
            assert (a["pred_species"], a["detected"]) == (b["pred_species"], b["detected"])
    assert (root / "outputs" / "detection_benchmark" / "summary.txt").exists()
    # 09: one detected and one undetected file, exit codes 0 / 2 and the reference's final line
    by = {}
    for r in rows:
        by.setdefault(r["detected"], r)
    for flag, r in by.items():
        wav = next((lse / "val_chunks").rglob(Path(r["file"]).name))
        with pytest.raises(SystemExit) as e:
            cli.main_09(["--wav", str(wav), "--config", str(cfgp), "--device", "cuda:0"], here=here)
        out = capsys.readouterr().out
        if str(flag) in ("1", "True", "true"):
            assert e.value.code == 0 and f"✅ DETECTADO: {r['pred_species']}" in out
        else:
            assert e.value.code == 2 and "❌ NO DETECTADO" in out
