"""TEST INFRASTRUCTURE ONLY -- golden artefacts of the reference's file-level pipeline.

Run in the build container (needs ``/root/reference``)::

    python -m oracle.make_golden_pipeline

Executes, unmodified and through their own ``main()`` (``sys.argv`` set as ``run_qout_grid.sh:13-59`` sets the flags):

    00_normalize_dataset_rms.process_folder     raw/{train,val}_chunks -> {train,val}_chunks     (00:41-57)
    for q_out in 0.10 0.15 0.20 0.25:                                                            (run_qout_grid.sh:13)
        08_fit_radial_detector.main   --root train_chunks --q-in 0.95 --q-out q --max-per-class 8 --cache
        10_benchmark_folder_detection.main --root val_chunks
        9105_make_config_snapshot_from_log.main

on a seeded synthetic WAV tree (``synth.write_wav_tree``: 4 species x 10 train / 6 val files of 3 s, re-creatable on the
GPU box).  Only three module attributes are replaced, none of them arithmetic: ``find_project_root`` (points at the temp
tree), ``load_encoder`` (returns the stand-in encoder: the thesis checkpoint is not public) and, in 10,
``load_eval_module`` (returns the already imported 09).  Everything the other reference scripts parse is stored under
``tests/golden/pipeline/``: per grid point ``config_used.json`` (radial_detector block), ``results.csv`` (paths made
relative), ``summary.txt``, ``run.log`` and ``config_snapshot.json``; plus sha256 of the normalised WAV files.
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import json
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

from oracle import ref_import  # noqa: E402
from amphibian_vae_latent_detector_b200 import synth  # noqa: E402
from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder  # noqa: E402

OUT = REPO / "tests" / "golden" / "pipeline"
SPECIES = ["Batrachyla_leptopus", "Batrachyla_taeniata", "Calyptocephalella_gayi", "Pleurodema_thaul"]
GRID = ("0.10", "0.15", "0.20", "0.25")
Q_IN = "0.95"
MAX_PER_CLASS = "8"
N_TRAIN, N_VAL, LENGTH = 10, 6, 144000
SEED_TRAIN, SEED_VAL = 4101, 4102


def make_raw_tree(root: Path) -> None:
    """The recipe the GPU test repeats (tests/test_gpu_pipeline.py)."""
    synth.write_wav_tree(root / "raw" / "train_chunks", SPECIES, N_TRAIN, LENGTH, seed=SEED_TRAIN, special_every=17)
    synth.write_wav_tree(root / "raw" / "val_chunks", SPECIES, N_VAL, LENGTH, seed=SEED_VAL, special_every=11)


def run_main(mod, argv):
    """``python <script> <flags>`` with stdout captured (what ``2>&1 | tee run.log`` keeps)."""
    buf = io.StringIO()
    old = sys.argv
    sys.argv = [getattr(mod, "__file__", "script")] + list(argv)
    try:
        with contextlib.redirect_stdout(buf):
            mod.main()
    finally:
        sys.argv = old
    return buf.getvalue()


def main() -> None:
    if not ref_import.available():
        raise SystemExit("reference tree not available; fixtures can only be made in the build container")
    ref00, ref08, ref09, ref10 = (ref_import.load(k) for k in ("00", "08", "09", "10"))
    ref9105 = ref_import._load_by_path("_ref_9105", ref_import.LSE / "9105_make_config_snapshot_from_log.py")
    encoder = build_standin_encoder(seed=123)
    if OUT.exists():
        shutil.rmtree(OUT)
    OUT.mkdir(parents=True)
    with tempfile.TemporaryDirectory() as td:
        root = Path(td).resolve()
        lse = root / "latent_space_exploration"
        lse.mkdir()
        mdir = root / "downloaded_models" / "bird_net_vae_audio_splitted_encoder_v0"
        mdir.mkdir(parents=True)
        (mdir / "model.pt").write_bytes(b"")                               # existence checks only: load_encoder is replaced
        (mdir / "bird_net_vae_audio_splitted.yaml").write_text("encoder: {}\n")
        make_raw_tree(lse)
        for split in ("train_chunks", "val_chunks"):
            ref00.process_folder(lse / "raw" / split, lse / split, sr=48000)
        shas = {}
        for split in ("train_chunks", "val_chunks"):
            for wav in sorted((lse / split).rglob("*.wav")):
                shas[str(wav.relative_to(lse))] = hashlib.sha256(wav.read_bytes()).hexdigest()
        cfg_path = root / "config.json"
        cfg_path.write_text(json.dumps({"species": SPECIES, "chunk_seconds": 3.0}, indent=2), encoding="utf-8")
        for mod in (ref08, ref10):
            mod.find_project_root = lambda start, _r=root: _r
        ref08.load_encoder = lambda *a, **k: encoder
        ref09.load_encoder = lambda *a, **k: encoder
        ref10.load_eval_module = lambda project_root: ref09
        for q in GRID:
            outdir = OUT / f"qout_{q}"
            outdir.mkdir()
            log = run_main(ref08, ["--config", str(cfg_path), "--root", str(lse / "train_chunks"), "--q-in", Q_IN,
                                   "--q-out", q, "--device", "cpu", "--max-per-class", MAX_PER_CLASS, "--cache"])
            log += run_main(ref10, ["--root", str(lse / "val_chunks"), "--config", str(cfg_path), "--device", "cpu"])
            log = log.replace(str(root), "<ROOT>")
            (outdir / "run.log").write_text(log, encoding="utf-8")
            bench = root / "outputs" / "detection_benchmark"
            (outdir / "summary.txt").write_text((bench / "summary.txt").read_text(encoding="utf-8"), encoding="utf-8")
            (outdir / "results.csv").write_text((bench / "results.csv").read_text(encoding="utf-8").replace(str(root), "<ROOT>"),
                                                encoding="utf-8")
            (outdir / "config_used.json").write_text(cfg_path.read_text(encoding="utf-8").replace(str(root), "<ROOT>"),
                                                     encoding="utf-8")
            tmp_log = root / "run.log"
            tmp_log.write_text(log, encoding="utf-8")
            run_main(ref9105, ["--log", str(tmp_log), "--q-in", Q_IN, "--q-out", q, "--out", str(root / "snap.json")])
            snap = json.loads((root / "snap.json").read_text())
            snap.pop("timestamp", None)
            snap["source_log"] = "run.log"
            (outdir / "config_snapshot.json").write_text(json.dumps(snap, indent=2), encoding="utf-8")
        cache = sorted(p.name for p in (lse / "cache_npz").glob("*.npz"))
        z0 = np.load(lse / "cache_npz" / cache[0])
        meta = {"species": SPECIES, "grid": GRID, "q_in": Q_IN, "max_per_class": MAX_PER_CLASS, "n_train": N_TRAIN,
                "n_val": N_VAL, "length": LENGTH, "seed_train": SEED_TRAIN, "seed_val": SEED_VAL, "numpy": np.__version__,
                "normalised_wav_sha256": shas, "cache_files": cache,
                "cache_keys": sorted(z0.files), "cache_Z_shape": list(z0["Z"].shape)}
        (OUT / "meta.json").write_text(json.dumps(meta, indent=1), encoding="utf-8")
    total = sum(p.stat().st_size for p in OUT.rglob("*") if p.is_file())
    print(f"wrote {total / 1e3:.1f} kB into {OUT}")


if __name__ == "__main__":
    main()
