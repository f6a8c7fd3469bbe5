"""Command-line surface of the hot path: the reference's scripts are driven as ``python NN_script.py <flags>`` by its bash glue
(``run_qout_grid.sh:28-38``, ``scripts/01…``, ``scripts/04…``), so the same flags, defaults, messages and exit codes are
accepted here and mapped onto the batched GPU pipeline (``pipeline.py``, ``reference_api.py``).  The launchers under
``latent_space_exploration/`` keep the reference's file names and only call the ``main_*`` functions below.

``--device`` is accepted for compatibility (the reference defaults to ``cpu`` and its glue passes it); there is no CPU path:
``cpu`` means "the current CUDA device", ``cuda:N`` selects GPU N.
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path
from typing import List, Optional, Sequence

from . import pipeline
from . import reference_api as api

# flag, type, default -- the mel / STFT flags every script shares (07:424-432, 08:348-354, 09:451-457, 10:316-322)
MEL_FLAGS = (("--sr", int, 48000), ("--n-mels", int, 64), ("--target-frames", int, 192), ("--fmin", float, 150.0),
             ("--fmax", float, 15000.0), ("--hop-length", int, 384), ("--n-fft", int, 2048))
ENCODER_DIR = ("downloaded_models", "bird_net_vae_audio_splitted_encoder_v0")


def _add(p: argparse.ArgumentParser, flags) -> None:
    for name, typ, default in flags:
        p.add_argument(name, type=typ, default=default)


def _mel_kw(a: argparse.Namespace) -> dict:
    return dict(sr=a.sr, n_mels=a.n_mels, target_frames=a.target_frames, fmin=a.fmin, fmax=a.fmax, hop_length=a.hop_length,
                n_fft=a.n_fft)


def find_project_root(start: Path) -> Path:
    """First ancestor (at most 15 levels up) holding ``downloaded_models/`` and ``latent_space_exploration/`` (08:69-75);
    ``start`` itself when there is none."""
    start = Path(start).resolve()
    for cand in [start, *start.parents][:16]:
        if (cand / "downloaded_models").exists() and (cand / "latent_space_exploration").exists():
            return cand
    return start


def _default_encoder_files(project_root: Path, pt: Optional[str], yml: Optional[str]):
    base = project_root.joinpath(*ENCODER_DIR)
    encoder_pt = Path(pt).resolve() if pt else base / "model.pt"
    encoder_yaml = Path(yml).resolve() if yml else base / "bird_net_vae_audio_splitted.yaml"
    if not encoder_pt.exists():
        raise SystemExit(f"❌ No encontré encoder .pt en: {encoder_pt}")          # 08:91-95
    if not encoder_yaml.exists():
        raise SystemExit(f"❌ No encontré encoder YAML en: {encoder_yaml}")      # 08:98-102
    return encoder_pt, encoder_yaml


def _device_note(device: str) -> None:
    if not str(device).startswith("cuda"):
        print(f"ℹ️ --device {device}: este backend solo corre en GPU (CUDA); se usa el dispositivo CUDA actual.")


def _resolve_root(arg: str, project_root: Path) -> Path:
    """A relative ``--root`` is tried against the cwd, the project root and ``<project>/latent_space_exploration``
    (08:394-420)."""
    root_in = Path(arg).expanduser()
    cands = [root_in] if root_in.is_absolute() else [Path.cwd() / root_in, project_root / root_in,
                                                     project_root / "latent_space_exploration" / root_in]
    for cand in cands:
        cand = cand.resolve()
        if cand.is_dir():
            return cand
    raise SystemExit("❌ No existe chunks_dir. Probé:\n" + "\n".join(f"   - {c.resolve()}" for c in cands))


# ---------------------------------------------------------------------------------------------- 00
def parser_00() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="RMS-normalise train/val/test chunk folders (00_normalize_dataset_rms.py)")
    p.add_argument("--base-dir", type=str, default="latent_space_exploration")
    p.add_argument("--sr", type=int, default=48000)
    return p


def main_00(argv: Optional[Sequence[str]] = None) -> None:
    a = parser_00().parse_args(argv)
    base = Path(a.base_dir).resolve()
    for s in ("train_chunks", "val_chunks", "test_chunks"):                      # 00:66-77
        src, dst = base / s, base / f"{s}_norm"
        if not src.exists():
            print(f"⚠ No existe {src}")
            continue
        print(f"\nProcesando {s} → {s}_norm")
        api.process_folder(src, dst, sr=a.sr)


# ---------------------------------------------------------------------------------------------- 07
def parser_07() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="WAV -> latent vector (07_encode_wav_to_latent.py)")
    p.add_argument("--wav", required=True, type=str)
    p.add_argument("--encoder", type=str, default=None)
    p.add_argument("--encoder-config", type=str, default=None)
    p.add_argument("--device", type=str, default="cpu")
    _add(p, MEL_FLAGS)
    p.add_argument("--duration", type=float, default=3.0)                       # 07:425 (the other scripts read config.json)
    p.add_argument("--auto-frames", action="store_true")
    p.add_argument("--auto-max-frames", type=int, default=512)
    p.add_argument("--auto-step", type=int, default=8)
    p.add_argument("--jsonl", action="store_true")
    p.add_argument("--precision", type=int, default=6)
    return p


def main_07(argv: Optional[Sequence[str]] = None, here: Optional[Path] = None) -> None:
    a = parser_07().parse_args(argv)
    project_root = find_project_root(here or Path.cwd())
    wav = Path(a.wav)
    wav = wav if wav.is_absolute() else (Path.cwd() / wav).resolve()
    base = project_root.joinpath(*ENCODER_DIR)
    encoder_path = Path(a.encoder) if a.encoder else base / "model.pt"
    if not encoder_path.exists():
        raise SystemExit("❌ No encontré encoder por defecto. Pasa la ruta con --encoder ...")       # 07:451-452
    yml = Path(a.encoder_config) if a.encoder_config else base / "bird_net_vae_audio_splitted.yaml"
    print(f"📌 Project root: {project_root}")
    print(f"🎧 WAV: {wav}")
    print(f"🧠 Encoder: {encoder_path}")
    if a.encoder_config:
        print(f"🧾 Encoder config: {yml}")
    print(f"🖥️ Device: {a.device}\n")
    _device_note(a.device)
    encoder = api.load_encoder(encoder_path, yml, project_root, None)
    api.encode_wav_report(wav, encoder, device="cpu", sr=a.sr, duration=a.duration, n_mels=a.n_mels, fmin=a.fmin, fmax=a.fmax,
                          hop_length=a.hop_length, n_fft=a.n_fft, target_frames=a.target_frames, auto_frames=a.auto_frames,
                          auto_max_frames=a.auto_max_frames, auto_step=a.auto_step, jsonl=a.jsonl, precision=a.precision)


# ---------------------------------------------------------------------------------------------- 08
def parser_08() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Fit the radial detector (08_fit_radial_detector.py)")
    p.add_argument("--config", type=str, default="config.json")
    p.add_argument("--root", type=str, required=True,
                   help="Carpeta con subcarpetas por especie (train_chunks/test_chunks/val_chunks)")
    p.add_argument("--q-in", type=float, default=0.95)
    p.add_argument("--q-out", type=float, default=0.01)
    p.add_argument("--device", type=str, default="cpu")
    _add(p, MEL_FLAGS)
    p.add_argument("--encoder-pt", type=str, default=None)
    p.add_argument("--encoder-yaml", type=str, default=None)
    p.add_argument("--max-per-class", type=int, default=0, help="0 = usar todos; si >0, samplea hasta este N por especie")
    p.add_argument("--seed", type=int, default=123)
    p.add_argument("--cache", action="store_true", help="Guardar/cargar latentes Z por especie en cache_npz/")
    return p


def main_08(argv: Optional[Sequence[str]] = None, here: Optional[Path] = None) -> None:
    a = parser_08().parse_args(argv)
    if not (0.0 < a.q_in < 1.0):
        raise SystemExit("❌ --q-in debe estar en (0,1).")
    if not (0.0 < a.q_out < 1.0):
        raise SystemExit("❌ --q-out debe estar en (0,1).")
    project_root = find_project_root(here or Path.cwd())
    cfg_path = Path(a.config)
    if not cfg_path.is_absolute():
        cfg_path = (project_root / cfg_path).resolve()
    if not cfg_path.exists():
        raise SystemExit(f"❌ No existe config.json en: {cfg_path}")
    cfg = api.load_json(cfg_path)
    species = cfg.get("species")
    if not isinstance(species, list) or not all(isinstance(s, str) for s in species):
        raise SystemExit("❌ config.json debe tener un campo 'species' (lista de strings).")
    chunks_dir = _resolve_root(a.root, project_root)
    encoder_pt, encoder_yaml = _default_encoder_files(project_root, a.encoder_pt, a.encoder_yaml)
    print(f"📌 Project root: {project_root}")
    print(f"🧾 Config: {cfg_path}")
    print(f"📁 Chunks dir: {chunks_dir}")
    print(f"🖥️ Device: {a.device}")
    _device_note(a.device)
    print("")
    print("📦 WAVs por especie en root:")
    for sp in species:
        d = chunks_dir / sp
        print(f"   - {sp}: {len(list(d.glob('*.wav'))) if d.exists() else 0}")
    print("")
    encoder = api.load_encoder(encoder_pt, encoder_yaml, project_root, None)
    pipeline.fit_radial_detector(cfg_path, chunks_dir, encoder, q_in=a.q_in, q_out=a.q_out, max_per_class=a.max_per_class,
                                 seed=a.seed, cache=a.cache,
                                 cache_dir=(project_root / "latent_space_exploration" / "cache_npz").resolve(),
                                 mel=_mel_kw(a))


# ---------------------------------------------------------------------------------------------- 09
def parser_09() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Detect the species of one WAV (09_evaluate_wav_detection.py)")
    p.add_argument("--wav", required=True, type=str, help="Ruta al archivo .wav a evaluar")
    p.add_argument("--config", type=str, default=None, help="Ruta a config.json (opcional)")
    p.add_argument("--encoder-pt", type=str, default=None, help="Ruta a model.pt (opcional)")
    p.add_argument("--encoder-yaml", type=str, default=None, help="Ruta a .yaml del encoder (opcional)")
    p.add_argument("--device", type=str, default="cpu")
    _add(p, MEL_FLAGS)
    return p


def main_09(argv: Optional[Sequence[str]] = None, here: Optional[Path] = None) -> None:
    a = parser_09().parse_args(argv)
    project_root = find_project_root(here or Path.cwd())
    config = a.config if a.config else str(project_root / "config.json")
    encoder_pt, encoder_yaml = _default_encoder_files(project_root, a.encoder_pt, a.encoder_yaml)
    _device_note(a.device)
    detected, sp = api.detect_species(a.wav, config_path=config, encoder_pt=str(encoder_pt), encoder_yaml=str(encoder_yaml),
                                      device=a.device, **_mel_kw(a))
    if detected:                                                                 # 09:478-483
        print(f"✅ DETECTADO: {sp}")
        sys.exit(0)
    print("❌ NO DETECTADO")
    sys.exit(2)


# ---------------------------------------------------------------------------------------------- 10
def parser_10() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Benchmark detection over a folder tree (10_benchmark_folder_detection.py)")
    p.add_argument("--root", type=str, default=None, help="Carpeta raíz a escanear (ej: latent_space_exploration/test_chunks)")
    p.add_argument("--config", type=str, default=None, help="Ruta a config.json (opcional)")
    p.add_argument("--encoder-pt", type=str, default=None, help="Ruta a model.pt del encoder (opcional)")
    p.add_argument("--encoder-yaml", type=str, default=None, help="Ruta a YAML del encoder (opcional)")
    p.add_argument("--device", type=str, default="cpu", help="cpu o cuda")
    _add(p, MEL_FLAGS)
    return p


def main_10(argv: Optional[Sequence[str]] = None, here: Optional[Path] = None) -> None:
    a = parser_10().parse_args(argv)
    project_root = find_project_root(here or Path.cwd())
    root = Path(a.root).expanduser().resolve() if a.root else project_root / "latent_space_exploration" / "test_chunks"
    if not root.exists():
        raise FileNotFoundError(f"No existe root: {root}")
    config_path = Path(a.config).expanduser().resolve() if a.config else project_root / "config.json"
    encoder_pt, encoder_yaml = _default_encoder_files(project_root, a.encoder_pt, a.encoder_yaml)
    out_dir = project_root / "outputs" / "detection_benchmark"                   # 10:345-346
    print("=" * 70)
    print("🔎 BENCHMARK DETECTION ON FOLDER")
    print(f"Root: {root}")
    print(f"Outputs: {out_dir}")
    print("=" * 70)
    _device_note(a.device)
    print("⏳ Cargando detector (config + encoder) una sola vez...")
    encoder = api.load_encoder(encoder_pt, encoder_yaml, project_root, None)
    print("✅ Listo.")
    pipeline.benchmark_folder(root, config_path, encoder, out_dir, mel=_mel_kw(a))
    print(f"\n✅ CSV guardado: {out_dir / 'results.csv'}")
    print(f"✅ Resumen guardado: {out_dir / 'summary.txt'}")


# ---------------------------------------------------------------------------------------------- 08b / 09n / 10b (MAP)
MAP_ENCODER_DIR = ("models", "bird_net_vae_audio_splitted_encoder_v0")          # core:64-77 (models/, not downloaded_models/)


def _map_default_files(project_root: Path, config: Optional[str], pt: Optional[str], yml: Optional[str],
                       config_relative_to_root: bool = False):
    """``resolve_default_config`` / ``_encoder_pt`` / ``_encoder_yaml`` of map_detector_core.py:56-77: explicit paths are
    taken as given, defaults must exist (``FileNotFoundError`` with the reference's wording)."""
    def must(p: Path, what: str) -> Path:
        if not p.exists():
            raise FileNotFoundError(f"No encontré {what} en: {p}")
        return p

    if config is None:
        cfg_path = must(project_root / "config.json", "config.json")
    else:
        cfg_path = Path(config)
        if config_relative_to_root:                                              # 08b:142-144
            cfg_path = cfg_path if cfg_path.is_absolute() else (project_root / cfg_path).resolve()
        else:                                                                    # 09n:88, 10b:315
            cfg_path = cfg_path.expanduser().resolve()
    base = project_root.joinpath(*MAP_ENCODER_DIR)
    encoder_pt = Path(pt).expanduser().resolve() if pt else must(base / "model.pt", "encoder .pt")
    encoder_yaml = Path(yml).expanduser().resolve() if yml else must(base / "bird_net_vae_audio_splitted.yaml", "encoder YAML")
    return cfg_path, encoder_pt, encoder_yaml


def parser_08b() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Fit the Gaussian-MAP detector (08b_fit_map_detector.py)")
    p.add_argument("--config", type=str, default="config.json")
    p.add_argument("--root", type=str, required=True, help="Carpeta con subcarpetas por especie (train_chunks/...)")
    p.add_argument("--device", type=str, default="cpu")
    _add(p, MEL_FLAGS)
    p.add_argument("--encoder-pt", type=str, default=None)
    p.add_argument("--encoder-yaml", type=str, default=None)
    p.add_argument("--max-per-class", type=int, default=0, help="0 = usar todos; si >0, samplea hasta este N por especie")
    p.add_argument("--seed", type=int, default=123)
    p.add_argument("--cache", action="store_true",
                   help="Guardar/cargar latentes Z por especie en latent_space_exploration/cache_npz/")
    p.add_argument("--cov-type", type=str, default="lda", choices=["lda", "qda"])
    p.add_argument("--cov-structure", type=str, default="full", choices=["full", "diag"])
    p.add_argument("--priors", type=str, default="empirical", choices=["empirical", "uniform"])
    p.add_argument("--eps", type=float, default=1e-6)
    p.add_argument("--shrink", type=float, default=0.0)
    p.add_argument("--set-tau-q", type=float, default=None, help="Ej: 0.01 => tau=quantile(scores_true,0.01)")
    return p


def main_08b(argv: Optional[Sequence[str]] = None, here: Optional[Path] = None) -> None:
    a = parser_08b().parse_args(argv)
    if not (0.0 <= a.shrink <= 1.0):
        raise SystemExit("❌ --shrink debe estar en [0,1].")                      # 08b:131-134
    if a.set_tau_q is not None and not (0.0 < float(a.set_tau_q) < 1.0):
        raise SystemExit("❌ --set-tau-q debe estar en (0,1).")
    project_root = find_project_root(here or Path.cwd())
    cfg_path = Path(a.config)
    cfg_path = cfg_path if cfg_path.is_absolute() else (project_root / cfg_path).resolve()
    cfg = api.load_json(cfg_path)
    species = cfg.get("species")
    if not isinstance(species, list) or not all(isinstance(s, str) for s in species):
        raise SystemExit("❌ config.json debe tener un campo 'species' (lista de strings).")
    chunks_dir = _resolve_root(a.root, project_root)
    _, encoder_pt, encoder_yaml = _map_default_files(project_root, str(cfg_path), a.encoder_pt, a.encoder_yaml)
    print(f"📌 Project root: {project_root}")
    print(f"🧾 Config: {cfg_path}")
    print(f"📁 Chunks dir: {chunks_dir}")
    print(f"🖥️ Device: {a.device}")
    _device_note(a.device)
    encoder = api.load_encoder(encoder_pt, encoder_yaml, project_root, None)
    pipeline.fit_map_detector(cfg_path, chunks_dir, encoder, cov_type=a.cov_type, cov_structure=a.cov_structure,
                              priors=a.priors, eps=a.eps, shrink=a.shrink, set_tau_q=a.set_tau_q,
                              max_per_class=a.max_per_class, seed=a.seed, cache=a.cache,
                              cache_dir=(project_root / "latent_space_exploration" / "cache_npz").resolve(), mel=_mel_kw(a))


def parser_09n() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="MAP detection of one WAV (09n_evaluate_wav_detection.py)")
    p.add_argument("--wav", required=True, type=str, help="Ruta al archivo .wav a evaluar")
    p.add_argument("--config", type=str, default=None, help="Ruta a config.json (opcional)")
    p.add_argument("--encoder-pt", type=str, default=None, help="Ruta a model.pt (opcional)")
    p.add_argument("--encoder-yaml", type=str, default=None, help="Ruta a .yaml del encoder (opcional)")
    p.add_argument("--device", type=str, default="cpu")
    _add(p, MEL_FLAGS)
    return p


def main_09n(argv: Optional[Sequence[str]] = None, here: Optional[Path] = None) -> None:
    a = parser_09n().parse_args(argv)
    project_root = find_project_root(here or Path.cwd())
    wav_p = Path(a.wav).expanduser()
    if not wav_p.is_absolute():
        wav_p = (Path.cwd() / wav_p).resolve()
    if not wav_p.exists():
        raise FileNotFoundError(f"No existe WAV: {wav_p}")                       # 09n:81-85, before anything is loaded
    cfg_path, encoder_pt, encoder_yaml = _map_default_files(project_root, a.config, a.encoder_pt, a.encoder_yaml)
    _device_note(a.device)
    detected, sp, best_score = api.detect_species_map(wav_p, config_path=str(cfg_path), encoder_pt=str(encoder_pt),
                                                      encoder_yaml=str(encoder_yaml), device=a.device, **_mel_kw(a))
    if detected:                                                                 # 09n:178-183
        print(f"✅ DETECTADO (MAP): {sp} | best_score={best_score:.6f}")
        sys.exit(0)
    print(f"❌ NO_DETECT (MAP) | best_score={best_score:.6f}")
    sys.exit(2)


def parser_10b() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="MAP benchmark over a folder tree (10b_benchmark_folder_detection_map.py)")
    p.add_argument("--root", type=str, default=None, help="Carpeta raíz a escanear (ej: latent_space_exploration/val_chunks)")
    p.add_argument("--config", type=str, default=None, help="Ruta a config.json (opcional)")
    p.add_argument("--encoder-pt", type=str, default=None, help="Ruta a model.pt (opcional)")
    p.add_argument("--encoder-yaml", type=str, default=None, help="Ruta a YAML del encoder (opcional)")
    p.add_argument("--device", type=str, default="cpu", help="cpu o cuda")
    _add(p, MEL_FLAGS)
    return p


def main_10b(argv: Optional[Sequence[str]] = None, here: Optional[Path] = None) -> None:
    project_root = find_project_root(here or Path.cwd())
    a = parser_10b().parse_args(argv)
    root = Path(a.root).expanduser().resolve() if a.root else project_root / "latent_space_exploration" / "val_chunks"
    if not root.exists():
        raise FileNotFoundError(f"No existe root: {root}")
    config_path, encoder_pt, encoder_yaml = _map_default_files(project_root, a.config, a.encoder_pt, a.encoder_yaml)
    out_dir = project_root / "outputs" / "detection_benchmark_map"              # 10b:318-319
    print("=" * 70)
    print("🔎 BENCHMARK DETECTION ON FOLDER — MAP")
    print(f"Root: {root}")
    print(f"Config: {config_path}")
    print(f"Outputs: {out_dir}")
    print("=" * 70)
    _device_note(a.device)
    print("⏳ Cargando detector MAP (config + encoder) una sola vez...")
    encoder = api.load_encoder(encoder_pt, encoder_yaml, project_root, None)
    print("✅ Listo.")
    pipeline.benchmark_folder_map(root, config_path, encoder, out_dir, mel=_mel_kw(a))
    print(f"\n✅ CSV guardado: {out_dir / 'results.csv'}")
    print(f"✅ Resumen guardado: {out_dir / 'summary.txt'}")


# ---------------------------------------------------------------------------------------------- q_out grid
def parser_grid() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="q_out grid in one process (run_qout_grid.sh: encode once, radii once)")
    p.add_argument("--config", type=str, default="config.json")
    p.add_argument("--train-root", type=str, default="train_chunks")
    p.add_argument("--val-root", type=str, default="val_chunks")
    p.add_argument("--grid-root", type=str, required=True)
    p.add_argument("--q-in", type=float, default=0.95)                           # run_qout_grid.sh:6
    p.add_argument("--q-out", type=float, nargs="+", default=[0.10, 0.15, 0.20, 0.25])      # :13
    p.add_argument("--max-per-class", type=int, default=400)                     # :8
    p.add_argument("--seed", type=int, default=123)
    p.add_argument("--device", type=str, default="cpu")
    p.add_argument("--encoder-pt", type=str, default=None)
    p.add_argument("--encoder-yaml", type=str, default=None)
    _add(p, MEL_FLAGS)
    return p


def main_grid(argv: Optional[Sequence[str]] = None, here: Optional[Path] = None) -> None:
    a = parser_grid().parse_args(argv)
    project_root = find_project_root(here or Path.cwd())
    cfg_path = Path(a.config)
    cfg_path = cfg_path if cfg_path.is_absolute() else (project_root / cfg_path).resolve()
    encoder_pt, encoder_yaml = _default_encoder_files(project_root, a.encoder_pt, a.encoder_yaml)
    _device_note(a.device)
    encoder = api.load_encoder(encoder_pt, encoder_yaml, project_root, None)
    out = pipeline.run_qout_grid(_resolve_root(a.train_root, project_root), _resolve_root(a.val_root, project_root), cfg_path,
                                 encoder, Path(a.grid_root), q_in=a.q_in, grid=list(a.q_out), max_per_class=a.max_per_class,
                                 seed=a.seed, mel=_mel_kw(a))
    for q, r in out.items():
        print(f"q_out={q}: Acc={r['acc'] * 100:.2f}% | NO_DETECT={r['no_detect'] * 100:.2f}% | rk={r['thresholds']}")


COMMANDS = {"normalize": main_00, "encode": main_07, "fit": main_08, "detect": main_09, "benchmark": main_10, "grid": main_grid,
            "fit-map": main_08b, "detect-map": main_09n, "benchmark-map": main_10b}


def main(argv: Optional[List[str]] = None) -> None:
    """``python -m amphibian_vae_latent_detector_b200.cli <normalize|encode|fit|detect|benchmark|grid|fit-map|detect-map|benchmark-map>
    <flags>``."""
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in COMMANDS:
        raise SystemExit("usage: cli.py {" + "|".join(COMMANDS) + "} <flags of the corresponding reference script>")
    COMMANDS[argv[0]](argv[1:])


if __name__ == "__main__":
    main()
