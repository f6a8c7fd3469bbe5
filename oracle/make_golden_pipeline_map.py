"""TEST INFRASTRUCTURE ONLY -- golden artefacts of the reference's file-level MAP pipeline.

Run in the build container (needs ``/root/reference``)::

    python -m oracle.make_golden_pipeline_map

Executes, unmodified and through their own ``main()`` / public function:

    00_normalize_dataset_rms.process_folder       raw/{train,val}_chunks -> {train,val}_chunks      (00:41-57)
    per case:
        08b_fit_map_detector.main                 --root train_chunks <case flags>                    (08b:128-358)
        10b_benchmark_folder_detection_map.main   --root val_chunks                                   (10b:306-407)
        09n_evaluate_wav_detection.detect_species_map on the first val file of every species          (09n:51-140)

on the seeded WAV tree of :mod:`oracle.make_golden_pipeline` (same recipe, re-creatable on the GPU box).  The three scripts
import ``latent_space_exploration.map_detector_core``; this repo ships a package of that name (the drop-in surface), so
the reference's core is bound to that name in ``sys.modules`` only while the scripts are loaded.  Replaced attributes, none
of them arithmetic: ``find_project_root`` (points at the temp tree) and ``load_encoder`` (returns the stand-in encoder).

Stored under ``tests/golden/pipeline_map/<case>/``: ``config_used.json`` with the three big matrices moved out
(``means`` / ``cov`` / ``precision`` -> ``params.npz``, float32 as the reference rounds them before writing),
``results.csv`` (paths made relative), ``summary.txt``, ``run.log``; plus ``meta.json`` (flags, 09n results).
"""
from __future__ import annotations

import importlib.util
import json
import shutil
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

from oracle import ref_import, shims  # noqa: E402
from oracle.make_golden_pipeline import SPECIES, make_raw_tree, run_main  # noqa: E402
from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder  # noqa: E402

OUT = REPO / "tests" / "golden" / "pipeline_map"
CASES = {
    "lda_diag_tau": ["--cov-type", "lda", "--cov-structure", "diag", "--shrink", "0.3", "--set-tau-q", "0.05",
                     "--max-per-class", "8", "--cache"],
    "qda_full_uniform": ["--cov-type", "qda", "--cov-structure", "full", "--shrink", "0.5", "--priors", "uniform"],
}


def load_map_scripts():
    """-> (ref 08b, ref 09n, ref 10b) with ``latent_space_exploration.map_detector_core`` = the reference's core."""
    shims.install()
    core = ref_import.load("core")
    try:
        import matplotlib  # noqa: F401
    except Exception:
        ref_import._install_matplotlib_stub()
    names = ("latent_space_exploration", "latent_space_exploration.map_detector_core")
    saved = {n: sys.modules.get(n) for n in names}
    pkg = types.ModuleType("latent_space_exploration")
    pkg.__path__ = [str(ref_import.LSE)]
    pkg.map_detector_core = core
    sys.modules[names[0]], sys.modules[names[1]] = pkg, core
    try:
        mods = []
        for tag, fn in (("08b", "08b_fit_map_detector.py"), ("09n", "09n_evaluate_wav_detection.py"),
                        ("10b", "10b_benchmark_folder_detection_map.py")):
            spec = importlib.util.spec_from_file_location(f"_ref_{tag}", str(ref_import.LSE / fn))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[spec.name] = mod                                           # 10b's @dataclass looks its module up
            spec.loader.exec_module(mod)
            mods.append(mod)
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
    return tuple(mods)


def split_params(cfg: dict, out_dir: Path) -> dict:
    """Move means / cov / precision of ``cfg['map_detector']`` into ``params.npz`` (species order = JSON order)."""
    md = cfg["map_detector"]
    names = list(md["means"])
    arrays = {k: np.stack([np.array(md[k][sp], dtype=np.float64) for sp in names]) for k in ("means", "cov", "precision")}
    for k, a in arrays.items():
        assert np.array_equal(a, a.astype(np.float32).astype(np.float64)), k      # the reference wrote float32 values
        md[k] = {sp: f"params.npz:{k}[{i}]" for i, sp in enumerate(names)}
    np.savez_compressed(out_dir / "params.npz", **{k: a.astype(np.float32) for k, a in arrays.items()})
    return cfg


def main() -> None:
    if not ref_import.available():
        raise SystemExit("reference tree not available; fixtures can only be made in the build container")
    ref00 = ref_import.load("00")
    ref08b, ref09n, ref10b = load_map_scripts()
    encoder = build_standin_encoder(seed=123)
    if OUT.exists():
        shutil.rmtree(OUT)
    OUT.mkdir(parents=True)
    meta = {"species": SPECIES, "cases": {}, "numpy": np.__version__}
    with tempfile.TemporaryDirectory() as td:
        root = Path(td).resolve()
        lse = root / "latent_space_exploration"
        lse.mkdir()
        (root / "downloaded_models").mkdir()
        mdir = root / "models" / "bird_net_vae_audio_splitted_encoder_v0"          # core:64-77
        mdir.mkdir(parents=True)
        (mdir / "model.pt").write_bytes(b"")                                        # existence checks only
        (mdir / "bird_net_vae_audio_splitted.yaml").write_text("encoder: {}\n")
        make_raw_tree(lse)
        for split in ("train_chunks", "val_chunks"):
            ref00.process_folder(lse / "raw" / split, lse / split, sr=48000)
        for mod in (ref08b, ref09n, ref10b):
            mod.find_project_root = lambda start, _r=root: _r
            mod.load_encoder = lambda *a, **k: encoder
        cfg_path = root / "config.json"
        for case, flags in CASES.items():
            outdir = OUT / case
            outdir.mkdir()
            cfg_path.write_text(json.dumps({"species": SPECIES, "chunk_seconds": 3.0}, indent=2), encoding="utf-8")
            log = run_main(ref08b, ["--config", str(cfg_path), "--root", str(lse / "train_chunks"), "--device", "cpu"] + flags)
            log += run_main(ref10b, ["--root", str(lse / "val_chunks"), "--config", str(cfg_path), "--device", "cpu"])
            (outdir / "run.log").write_text(log.replace(str(root), "<ROOT>"), encoding="utf-8")
            bench = root / "outputs" / "detection_benchmark_map"
            (outdir / "summary.txt").write_text((bench / "summary.txt").read_text(encoding="utf-8"), encoding="utf-8")
            (outdir / "results.csv").write_text((bench / "results.csv").read_text(encoding="utf-8").replace(str(root), "<ROOT>"),
                                                encoding="utf-8")
            cfg = json.loads(cfg_path.read_text(encoding="utf-8").replace(str(root), "<ROOT>"))
            (outdir / "config_used.json").write_text(json.dumps(split_params(cfg, outdir), indent=2, ensure_ascii=False),
                                                     encoding="utf-8")
            single = {}
            for sp in SPECIES:
                wav = sorted((lse / "val_chunks" / sp).glob("*.wav"))[0]
                det, name, best = ref09n.detect_species_map(wav, config_path=cfg_path, device="cpu")
                single[str(wav.relative_to(lse))] = [bool(det), name, float(best)]
            meta["cases"][case] = {"flags": flags, "detect_species_map": single}
    (OUT / "meta.json").write_text(json.dumps(meta, indent=1), encoding="utf-8")
    total = sum(p.stat().st_size for p in OUT.rglob("*") if p.is_file())
    print(f"wrote {total / 1e3:.1f} kB into {OUT}")


if __name__ == "__main__":
    main()
