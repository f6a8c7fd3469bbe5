"""Callers and on-disk formats either side of the hot path: what
``08_fit_radial_detector.py``, ``10_benchmark_folder_detection.py`` and ``run_qout_grid.sh`` do around it, with the
per-file loops replaced by GPU batches.  Every artefact keeps the reference's format because other reference scripts
parse them:

* stdout line ``✅ <sp>: rk_in=… | rk_out=… | rk=…`` (08:556) -> ``9105_make_config_snapshot_from_log.py:11-13``
* ``config.json``: ``radial_detector = {centroids, thresholds, meta_fit{…, per_species}}`` (08:561-583), ``.bak`` copy (:585-586)
* ``cache_npz/Z_<chunks_dir.name>_<species>.npz`` with arrays ``Z``, ``failed``, ``root`` (08:467-475, :518-520)
* ``results.csv`` (10:401-428) and ``summary.txt`` (10:278-301) -> ``9100_spearman_rk_analysis.py:53-58``
* per grid point ``qout_<q>/{run.log, summary.txt, results.csv, config_used.json, config_snapshot.json}``
  (run_qout_grid.sh:13-59; snapshot schema 9105:50-58)

* MAP detector (row N1): ``config.json`` ``map_detector = {model, cov_type, cov_structure, priors, means, cov, precision,
  logdet_cov, tau, meta_fit{…, score_true_global_summary, per_species}}`` (08b:322-351) and the MAP benchmark's ``results.csv``
  (column ``best_score``) / ``summary.txt`` (``… Summary (MAP)``) under ``outputs/detection_benchmark_map`` (10b:236-239, :357-385)

Plots (matplotlib) are out of scope.
"""
from __future__ import annotations

import json
import random
import shutil
from datetime import datetime
from pathlib import Path
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import reference_api as api

MEL_DEFAULTS = dict(sr=48000, n_mels=64, target_frames=192, fmin=150.0, fmax=15000.0, hop_length=384, n_fft=2048)


def _load_cfg(cfg_path: Path) -> Dict[str, Any]:
    cfg = api.load_json(cfg_path)
    species = cfg.get("species")
    if not isinstance(species, list) or not all(isinstance(s, str) for s in species):
        raise SystemExit("❌ config.json debe contener 'species' como lista de strings.")          # 08:388-390
    return cfg


def _chunk_seconds(cfg: Dict[str, Any]) -> float:
    try:
        return float(cfg.get("chunk_seconds", 5.0))                                               # 08:392-396
    except Exception:
        return 5.0


def encode_species_folders(encoder, chunks_dir: Path, species_list: Sequence[str], chunk_seconds: float, *,
                           max_per_class: int = 0, cache_dir: Optional[Path] = None, mel: Optional[dict] = None,
                           log: Callable[[str], None] = print, style: str = "08"):
    """08:461-520 (= 08b:186-243): per species, cache hit or ``sorted(glob)`` (+ ``random.sample``) -> batched encode ->
    optional cache.  ``style`` picks the wording of the two skip messages that differ between 08 and 08b."""
    mel = {**MEL_DEFAULTS, **(mel or {})}
    Z_by, failed_by, used_by = {}, {}, {}
    for sp in species_list:
        sp_dir = (chunks_dir / sp).resolve()
        if not sp_dir.exists():
            log(f"⚠️ {sp}: carpeta no existe (se omite): {sp_dir}" if style == "08b" else
                f"⚠️ {sp}: carpeta no existe: {sp_dir} (se omite).")
            continue
        cache_path = (cache_dir / f"Z_{chunks_dir.name}_{sp}.npz") if cache_dir is not None else None
        if cache_path is not None and cache_path.exists():
            data = np.load(cache_path)
            Zm = data["Z"].astype(np.float32)
            Z_by[sp], failed_by[sp], used_by[sp] = Zm, (int(data["failed"]) if "failed" in data else 0), int(Zm.shape[0])
            log(f"🧊 {sp}: cargado cache {cache_path.name} -> N={Zm.shape[0]}")
            continue
        wavs = sorted(sp_dir.glob("*.wav"))
        if len(wavs) == 0:
            log(f"⚠️ {sp}: sin wavs (se omite)." if style == "08b" else f"⚠️ {sp}: sin wavs en {sp_dir} (se omite).")
            continue
        if max_per_class and len(wavs) > max_per_class:
            wavs = random.sample(wavs, max_per_class)          # Python MT19937, state carried across species (08:483-484)
        Zm, failed = api.encode_wavs_to_latents(encoder, wavs, None, duration=chunk_seconds, return_failed=True, **mel)
        for w in failed:
            log(f"⚠️ {sp}: fallo {Path(w).name}")
        if Zm.shape[0] == 0:
            log(f"❌ {sp}: no se pudo codificar nada (se omite).")
            continue
        Z_by[sp], failed_by[sp], used_by[sp] = Zm.astype(np.float32), len(failed), int(Zm.shape[0])
        log(f"🧪 {sp}: encoded N={Zm.shape[0]} (failed={len(failed)})")
        if cache_path is not None:
            cache_path.parent.mkdir(parents=True, exist_ok=True)
            np.savez_compressed(cache_path, Z=Z_by[sp], failed=len(failed), root=str(chunks_dir))
            log(f"   ↳ guardado cache: {cache_path.name}")
    if not Z_by:
        raise SystemExit("❌ No se codificó ninguna especie. Revisa root y/o pipeline.")
    return Z_by, failed_by, used_by


def _fit_grid(Z_by: Dict[str, np.ndarray], q_in: float, q_outs: Sequence[float]):
    """One GPU pass for every species and every q_out of the grid (radii are q independent)."""
    eng = api._engine(144000, 0)
    names = list(Z_by.keys())
    Z = torch.from_numpy(np.concatenate([Z_by[sp] for sp in names])).to(eng.device)
    lab = torch.from_numpy(np.concatenate([np.full(Z_by[sp].shape[0], i, np.int32) for i, sp in enumerate(names)])).to(eng.device)
    return names, eng.fit_radial(Z, lab, len(names), q_in, list(q_outs))


def _summary(vals: np.ndarray) -> Dict[str, float]:
    return {"min": float(vals[0]), "p50": float(vals[1]), "p90": float(vals[2]), "max": float(vals[3])}


def radial_config_block(names, fit, qi: int, Z_by, failed_by, used_by, chunks_dir: Path, chunk_seconds: float, *, q_in,
                        q_out, max_per_class, seed, mel, log: Callable[[str], None] = print) -> Dict[str, Any]:
    """centroids / thresholds / meta_fit for grid point ``qi`` and the stdout lines of 08:556-558."""
    centroids, thresholds, meta = {}, {}, {}
    n_tot = int(sum(Z_by[sp].shape[0] for sp in names))
    for k, sp in enumerate(names):
        rk_in, rk_out, rk = float(fit.rk_in[k]), float(fit.rk_out[qi, k]), float(fit.rk[qi, k])
        n_in = int(Z_by[sp].shape[0])
        extra = {"rho_in_summary": _summary(fit.summaries["in"][k]),
                 "rho_out_summary": _summary(fit.summaries["out"][k]) if n_tot > n_in else
                 {"min": float("nan"), "p50": float("nan"), "p90": float("nan"), "max": float("nan")}}
        centroids[sp] = [float(v) for v in fit.centroids[k]]
        thresholds[sp] = rk
        meta[sp] = {"N_in": n_in, "N_out": n_tot - n_in, "rk_in": rk_in, "rk_out": rk_out if np.isfinite(rk_out) else None,
                    "rk_final": rk, "failed": int(failed_by.get(sp, 0)), "used": int(used_by.get(sp, n_in)), **extra}
        rk_out_print = rk_out if np.isfinite(rk_out) else float("nan")
        log(f"✅ {sp}: rk_in={rk_in:.6f} | rk_out={rk_out_print:.6f} | rk={rk:.6f}")
        log(f"   rho_in:  {extra['rho_in_summary']}")
        log(f"   rho_out: {extra['rho_out_summary']}")
    return {"centroids": centroids, "thresholds": thresholds,
            "meta_fit": {"chunks_dir": str(chunks_dir), "chunks_name": chunks_dir.name, "q_in": float(q_in),
                         "q_out": float(q_out), "chunk_seconds": float(chunk_seconds), "sr": int(mel["sr"]),
                         "n_mels": int(mel["n_mels"]), "target_frames": int(mel["target_frames"]), "fmin": float(mel["fmin"]),
                         "fmax": float(mel["fmax"]), "hop_length": int(mel["hop_length"]), "n_fft": int(mel["n_fft"]),
                         "max_per_class": int(max_per_class), "seed": int(seed), "per_species": meta}}


def fit_radial_detector(config_path: Path, root: Path, encoder, *, q_in: float = 0.95, q_out: float = 0.01,
                        max_per_class: int = 0, seed: int = 123, cache: bool = False, cache_dir: Optional[Path] = None,
                        mel: Optional[dict] = None, log: Callable[[str], None] = print) -> Dict[str, Any]:
    """``08_fit_radial_detector.py`` main (08:365-587) with batched GPU work; rewrites ``config.json`` (+ ``.bak``)."""
    if not (0.0 < q_in < 1.0):
        raise SystemExit("❌ --q-in debe estar en (0,1).")
    if not (0.0 < q_out < 1.0):
        raise SystemExit("❌ --q-out debe estar en (0,1).")
    random.seed(seed)
    np.random.seed(seed)
    cfg_path, chunks_dir = Path(config_path).resolve(), Path(root).resolve()
    cfg = _load_cfg(cfg_path)
    mel_kw = {**MEL_DEFAULTS, **(mel or {})}
    chunk_seconds = _chunk_seconds(cfg)
    log(f"🎯 q_in={q_in} | q_out={q_out} | max_per_class={max_per_class} | cache={cache}")
    cdir = (Path(cache_dir) if cache_dir else cfg_path.parent / "latent_space_exploration" / "cache_npz") if cache else None
    Z_by, failed_by, used_by = encode_species_folders(encoder, chunks_dir, cfg["species"], chunk_seconds,
                                                      max_per_class=max_per_class, cache_dir=cdir, mel=mel_kw, log=log)
    names, fit = _fit_grid(Z_by, q_in, [q_out])
    block = radial_config_block(names, fit, 0, Z_by, failed_by, used_by, chunks_dir, chunk_seconds, q_in=q_in, q_out=q_out,
                                max_per_class=max_per_class, seed=seed, mel=mel_kw, log=log)
    if not isinstance(cfg.get("radial_detector"), dict):
        cfg["radial_detector"] = {}
    cfg["radial_detector"].update(block)
    backup = cfg_path.with_suffix(cfg_path.suffix + ".bak")
    shutil.copy2(cfg_path, backup)
    cfg_path.write_text(json.dumps(cfg, indent=2, ensure_ascii=False), encoding="utf-8")
    log(f"\n💾 Guardado en: {cfg_path}")
    log(f"🗂️ Backup: {backup}")
    return cfg


def _list_audio_files(root: Path) -> List[Path]:
    return sorted(p for ext in (".wav", ".WAV") for p in root.rglob(f"*{ext}"))                   # 10:98-103


def write_summary(rows: List[Dict[str, Any]], out_txt: Path, title: str = "=== Detection Benchmark Summary ===") -> None:
    """10:278-301 / 10b:257-283 (pandas groupby / sort, so the per-class order matches the reference's)."""
    import pandas as pd
    df = pd.DataFrame(rows)
    total = len(df)
    correct = int(df["correct"].sum()) if total else 0
    acc = (correct / total) if total else 0.0
    no_det = int((df["pred_species"] == "NO_DETECT").sum()) if total else 0
    no_det_rate = (no_det / total) if total else 0.0
    lines = [title, f"Total files: {total}",
             f"Correct: {correct}  | Accuracy: {acc*100:.2f}%", f"NO_DETECT: {no_det} | Rate: {no_det_rate*100:.2f}%", "",
             "=== Per-class ==="]
    if total:
        per_class = df.groupby("true_species").agg(
            n=("file", "count"), acc=("correct", "mean"),
            no_detect=("pred_species", lambda s: (s == "NO_DETECT").mean())).sort_values("acc", ascending=False)
        for sp, row in per_class.iterrows():
            lines.append(f"- {sp:30s}  n={int(row['n']):4d}  acc={row['acc']*100:6.2f}%  no_detect={row['no_detect']*100:6.2f}%")
    Path(out_txt).write_text("\n".join(lines), encoding="utf-8")


def _results_rows(files, true_sp, results, value_col: str = "best_distance"):
    """One CSV row per file; the per-file number is ``best_distance`` for the radial detector (10:401-418) and
    ``best_score`` for MAP (10b:340-355)."""
    rows = []
    for f, t, (det, pred, best) in zip(files, true_sp, results):
        if pred == "ERROR":
            rows.append({"file": str(f), "true_species": t, "pred_species": "ERROR", "detected": False, "correct": False,
                         value_col: np.nan, "error": "unreadable file"})
        else:
            ps = pred if det and pred is not None else "NO_DETECT"
            rows.append({"file": str(f), "true_species": t, "pred_species": ps, "detected": bool(det),
                         "correct": bool(ps == t), value_col: float(best)})
    return rows


def _write_results(rows, out_dir: Path, title: str = "=== Detection Benchmark Summary ===") -> None:
    import pandas as pd
    out_dir.mkdir(parents=True, exist_ok=True)
    pd.DataFrame(rows).to_csv(out_dir / "results.csv", index=False, encoding="utf-8")            # 10:423-428
    write_summary([r for r in rows if r["pred_species"] != "ERROR"], out_dir / "summary.txt", title)     # 10:432


def benchmark_folder(root: Path, config_path: Path, encoder, out_dir: Path, *, mel: Optional[dict] = None,
                     log: Callable[[str], None] = print) -> List[Dict[str, Any]]:
    """``10_benchmark_folder_detection.py`` main (10:326-456): ``<root>/<true_species>/**.wav`` -> results.csv + summary.txt."""
    root = Path(root).resolve()
    cfg = api.load_json(Path(config_path))
    sess = api.DetectorSession(None, Path(config_path).parent, Path(config_path), Path("unused.pt"), Path("unused.yaml"),
                               "cuda", **{**MEL_DEFAULTS, **(mel or {})})
    sess.centroids, sess.thresholds, sess.duration = api.get_detector_from_config(cfg)
    sess.encoder = encoder
    class_dirs = [p for p in root.iterdir() if p.is_dir()]
    if not class_dirs:
        raise RuntimeError(f"No hay subcarpetas en root: {root}")
    files, true_sp = [], []
    for class_dir in sorted(class_dirs):
        wavs = _list_audio_files(class_dir)
        if not wavs:
            log(f"⚠️ Sin wavs en {class_dir}")
            continue
        log(f"\n📁 {class_dir.name}: {len(wavs)} archivos")
        files += wavs
        true_sp += [class_dir.name] * len(wavs)
    if not files:
        raise RuntimeError("No se procesó ningún archivo (rows vacío).")
    rows = _results_rows(files, true_sp, sess.predict_many(files))
    _write_results(rows, Path(out_dir))
    ok = [r for r in rows if r["pred_species"] != "ERROR"]
    acc = float(np.mean([r["correct"] for r in ok])) if ok else 0.0
    nd = float(np.mean([r["pred_species"] == "NO_DETECT" for r in ok])) if ok else 0.0
    log("\n" + "=" * 70)
    log(f"✅ DONE  | N={len(ok)} | Acc={acc*100:.2f}% | NO_DETECT={nd*100:.2f}%")
    log("=" * 70)
    return rows


def run_qout_grid(train_root: Path, val_root: Path, config_path: Path, encoder, grid_root: Path, *, q_in: float = 0.95,
                  grid: Sequence[float] = (0.10, 0.15, 0.20, 0.25), max_per_class: int = 400, seed: int = 123,
                  mel: Optional[dict] = None) -> Dict[str, Any]:
    """``run_qout_grid.sh`` (:13-59) in one process: train and val latents are encoded ONCE, the radii of every train
    latent to every centroid once, then per q_out only the quantile read-out, the decisions on the cached val latents
    and the files change.  Writes ``<grid_root>/qout_<q>/`` exactly as the script does (minus the PNGs)."""
    random.seed(seed)
    np.random.seed(seed)
    cfg_path = Path(config_path).resolve()
    cfg = _load_cfg(cfg_path)
    mel_kw = {**MEL_DEFAULTS, **(mel or {})}
    chunk_seconds = _chunk_seconds(cfg)
    train_root, val_root, grid_root = Path(train_root).resolve(), Path(val_root).resolve(), Path(grid_root)
    prelude: List[str] = []
    Z_by, failed_by, used_by = encode_species_folders(encoder, train_root, cfg["species"], chunk_seconds,
                                                      max_per_class=max_per_class, mel=mel_kw, log=prelude.append)
    names, fit = _fit_grid(Z_by, q_in, list(grid))
    # validation latents, once
    files, true_sp = [], []
    for class_dir in sorted(p for p in val_root.iterdir() if p.is_dir()):
        wavs = _list_audio_files(class_dir)
        files += wavs
        true_sp += [class_dir.name] * len(wavs)
    Zval, failed = api.encode_wavs_to_latents(encoder, files, None, duration=chunk_seconds, return_failed=True, **mel_kw)
    bad = set(map(str, failed))
    out: Dict[str, Any] = {}
    for qi, q in enumerate(grid):
        outdir = grid_root / f"qout_{q:.2f}"
        outdir.mkdir(parents=True, exist_ok=True)
        lines = list(prelude)
        block = radial_config_block(names, fit, qi, Z_by, failed_by, used_by, train_root, chunk_seconds, q_in=q_in, q_out=q,
                                    max_per_class=max_per_class, seed=seed, mel=mel_kw, log=lines.append)
        cfg_q = dict(cfg)
        cfg_q["radial_detector"] = block
        cents = {sp: np.array(v, dtype=np.float32) for sp, v in block["centroids"].items()}
        good = iter(api._decide_many(Zval, cents, block["thresholds"]))
        results = [(False, "ERROR", float("nan")) if str(f) in bad else next(good) for f in files]
        rows = _results_rows(files, true_sp, results)
        _write_results(rows, outdir)
        (outdir / "run.log").write_text("\n".join(lines) + "\n", encoding="utf-8")
        (outdir / "config_used.json").write_text(json.dumps(cfg_q, indent=2, ensure_ascii=False), encoding="utf-8")
        snapshot = {"timestamp": datetime.now().isoformat(), "q_in": float(q_in), "q_out": float(q),
                    "rk_in_per_species": {sp: round(float(fit.rk_in[k]), 6) for k, sp in enumerate(names)},
                    "rk_out_per_species": {sp: round(float(fit.rk_out[qi, k]), 6) for k, sp in enumerate(names)},
                    "rk_per_species": {sp: round(float(fit.rk[qi, k]), 6) for k, sp in enumerate(names)},
                    "source_log": str(outdir / "run.log")}                                        # 9105:50-58
        (outdir / "config_snapshot.json").write_text(json.dumps(snapshot, indent=2), encoding="utf-8")
        ok = [r for r in rows if r["pred_species"] != "ERROR"]
        out[f"{q:.2f}"] = {"acc": float(np.mean([r["correct"] for r in ok])) if ok else 0.0,
                           "no_detect": float(np.mean([r["pred_species"] == "NO_DETECT" for r in ok])) if ok else 0.0,
                           "thresholds": block["thresholds"]}
    # the config on disk ends up holding the last grid point, as after the shell loop
    shutil.copy2(cfg_path, cfg_path.with_suffix(cfg_path.suffix + ".bak"))
    cfg["radial_detector"] = block
    cfg_path.write_text(json.dumps(cfg, indent=2, ensure_ascii=False), encoding="utf-8")
    return out


# ----------------------------------------------------------------------------------------------------------
# Row N1 at file level: 08b_fit_map_detector.py and 10b_benchmark_folder_detection_map.py
# ----------------------------------------------------------------------------------------------------------
def summarize_1d(x: np.ndarray) -> Dict[str, float]:
    """core:92-101: min / p05 / p50 / p95 / max, NaN for an empty array."""
    x = np.asarray(x)
    if x.size == 0:
        return dict.fromkeys(("min", "p05", "p50", "p95", "max"), float("nan"))
    return {"min": float(np.min(x)), "p05": float(np.quantile(x, 0.05)), "p50": float(np.quantile(x, 0.50)),
            "p95": float(np.quantile(x, 0.95)), "max": float(np.max(x))}


def map_config_block(fit, scores_true_by: Dict[str, np.ndarray], failed_by, used_by, chunks_dir: Path, chunk_seconds: float,
                     *, cov_type, cov_structure, priors, eps, shrink, set_tau_q, max_per_class, seed, mel) -> Dict[str, Any]:
    """The ``map_detector`` object of 08b:322-351 (same keys, same order, same Python types) from a :class:`MapFit` and the
    float64 score of every training latent under its own class (08b:298-319)."""
    names = list(fit.species)
    per_species = {}
    for k, sp in enumerate(names):
        s = np.asarray(scores_true_by[sp], dtype=np.float64)
        per_species[sp] = {"N": int(fit.counts[k]), "failed": int(failed_by.get(sp, 0)),
                           "used": int(used_by.get(sp, fit.counts[k])), "prior": float(fit.priors[k]),
                           "score_true_summary": summarize_1d(s.astype(np.float32))}
    all_scores = np.concatenate([np.asarray(scores_true_by[sp], dtype=np.float64) for sp in names]) if names else np.zeros(0)
    return {
        "model": "gaussian_map", "cov_type": str(cov_type), "cov_structure": str(cov_structure), "priors": str(priors),
        "means": {sp: fit.means[k].astype(float).tolist() for k, sp in enumerate(names)},
        "cov": {sp: fit.cov[k].astype(float).tolist() for k, sp in enumerate(names)},
        "precision": {sp: fit.precision[k].astype(float).tolist() for k, sp in enumerate(names)},
        "logdet_cov": {sp: float(fit.logdet_cov[k]) for k, sp in enumerate(names)},
        "tau": fit.tau,
        "meta_fit": {"chunks_dir": str(chunks_dir), "chunks_name": chunks_dir.name, "chunk_seconds": float(chunk_seconds),
                     "sr": int(mel["sr"]), "n_mels": int(mel["n_mels"]), "target_frames": int(mel["target_frames"]),
                     "fmin": float(mel["fmin"]), "fmax": float(mel["fmax"]), "hop_length": int(mel["hop_length"]),
                     "n_fft": int(mel["n_fft"]), "max_per_class": int(max_per_class), "seed": int(seed), "eps": float(eps),
                     "shrink": float(shrink),
                     "tau_from_train_quantile": float(set_tau_q) if set_tau_q is not None else None,
                     "score_true_global_summary": summarize_1d(all_scores.astype(np.float32)),
                     "per_species": per_species}}


def fit_map_detector(config_path: Path, root: Path, encoder, *, cov_type: str = "lda", cov_structure: str = "full",
                     priors: str = "empirical", eps: float = 1e-6, shrink: float = 0.0, set_tau_q: Optional[float] = None,
                     max_per_class: int = 0, seed: int = 123, cache: bool = False, cache_dir: Optional[Path] = None,
                     mel: Optional[dict] = None, log: Callable[[str], None] = print) -> Dict[str, Any]:
    """``08b_fit_map_detector.py`` main (08b:128-358): encode the species folders in GPU batches, fit the Gaussian-MAP model
    (means and second moments on the device, D x D algebra on the host), score every training latent under its own class,
    optionally set ``tau`` to a quantile of those scores, and rewrite ``config.json`` (+ ``.bak``)."""
    if not (0.0 <= shrink <= 1.0):
        raise SystemExit("❌ --shrink debe estar en [0,1].")
    if set_tau_q is not None and not (0.0 < float(set_tau_q) < 1.0):
        raise SystemExit("❌ --set-tau-q debe estar en (0,1).")
    random.seed(seed)
    np.random.seed(seed)
    cfg_path, chunks_dir = Path(config_path).resolve(), Path(root).resolve()
    cfg = api.load_json(cfg_path)
    species_list = cfg.get("species")
    if not isinstance(species_list, list) or not all(isinstance(s, str) for s in species_list):
        raise SystemExit("❌ config.json debe tener un campo 'species' (lista de strings).")       # 08b:148-150
    mel_kw = {**MEL_DEFAULTS, **(mel or {})}
    chunk_seconds = _chunk_seconds(cfg)
    log(f"🎯 cov_type={cov_type} | cov_structure={cov_structure} | priors={priors} | eps={eps} | shrink={shrink}")
    log(f"🎯 max_per_class={max_per_class} | cache={cache}\n")
    cdir = (Path(cache_dir) if cache_dir else cfg_path.parent / "latent_space_exploration" / "cache_npz") if cache else None
    Z_by, failed_by, used_by = encode_species_folders(encoder, chunks_dir, species_list, chunk_seconds,
                                                      max_per_class=max_per_class, cache_dir=cdir, mel=mel_kw, log=log,
                                                      style="08b")
    names = sorted(Z_by)                                                                          # 08b:248
    eng = api._engine(144000, 0)
    Z = torch.from_numpy(np.concatenate([Z_by[sp] for sp in names])).to(eng.device)
    lab_np = np.concatenate([np.full(Z_by[sp].shape[0], i, np.int32) for i, sp in enumerate(names)])
    lab = torch.from_numpy(lab_np).to(eng.device)
    fit = eng.fit_map(Z, lab, names, cov_type=cov_type, cov_structure=cov_structure, priors=priors, eps=float(eps),
                      shrink=float(shrink))
    _, _, scores = eng.map_score(Z, fit, want_scores=True, tau=None)
    scores = scores.cpu().numpy()
    true_scores = scores[np.arange(scores.shape[0]), lab_np]
    scores_true_by = {sp: true_scores[lab_np == i] for i, sp in enumerate(names)}
    if set_tau_q is not None:
        fit.tau = float(np.quantile(true_scores.astype(np.float64), float(set_tau_q)))            # 08b:315-319
        log(f"\n✅ tau fijado desde train: tau = quantile(score_true_class, q={float(set_tau_q)}) = {fit.tau:.6f}")
    cfg["map_detector"] = map_config_block(fit, scores_true_by, failed_by, used_by, chunks_dir, chunk_seconds,
                                           cov_type=cov_type, cov_structure=cov_structure, priors=priors, eps=eps,
                                           shrink=shrink, set_tau_q=set_tau_q, max_per_class=max_per_class, seed=seed,
                                           mel=mel_kw)
    backup = cfg_path.with_suffix(cfg_path.suffix + ".bak")
    shutil.copy2(cfg_path, backup)
    cfg_path.write_text(json.dumps(cfg, indent=2, ensure_ascii=False), encoding="utf-8")
    log(f"\n💾 Guardado en: {cfg_path}")
    log(f"🗂️ Backup: {backup}")
    log("\n✅ MAP detector fit listo. (NO_DETECT se decide con tau en 09n/10b.)")
    return cfg


def benchmark_folder_map(root: Path, config_path: Path, encoder, out_dir: Path, *, mel: Optional[dict] = None,
                         log: Callable[[str], None] = print) -> List[Dict[str, Any]]:
    """``10b_benchmark_folder_detection_map.py`` main (10b:306-407): ``<root>/<true_species>/**.wav`` scored by the MAP
    detector of ``config.json`` in GPU batches -> ``results.csv`` (``best_score``) + ``summary.txt``."""
    root = Path(root).resolve()
    cfg = api.load_json(Path(config_path))
    sess = api.MapDetectorSession(Path(config_path).parent, Path(config_path), Path("unused.pt"), Path("unused.yaml"), "cuda",
                                  **{**MEL_DEFAULTS, **(mel or {})})
    sess.set_params(cfg)
    sess.encoder = encoder
    log(f"⏱️ chunk_seconds usados: {sess.duration}")
    log(f"🎯 tau (rechazo): {sess.tau}\n")
    class_dirs = [d for d in root.iterdir() if d.is_dir() and not d.name.startswith(".")]          # 10b:332
    if not class_dirs:
        raise RuntimeError(f"No encontré subcarpetas de especies en: {root}")
    files, true_sp = [], []
    for class_dir in sorted(class_dirs):
        wavs = _list_audio_files(class_dir)
        if not wavs:
            log(f"⚠️ Sin wavs en {class_dir}")
            continue
        log(f"\n📁 {class_dir.name}: {len(wavs)} archivos")
        files += wavs
        true_sp += [class_dir.name] * len(wavs)
    if not files:
        raise RuntimeError("No se procesó ningún archivo (rows vacío).")
    rows = _results_rows(files, true_sp, sess.predict_many(files), value_col="best_score")
    _write_results(rows, Path(out_dir), title="=== Detection Benchmark Summary (MAP) ===")
    ok = [r for r in rows if r["pred_species"] != "ERROR"]
    acc = float(np.mean([r["correct"] for r in ok])) if ok else 0.0
    nd = float(np.mean([r["pred_species"] == "NO_DETECT" for r in ok])) if ok else 0.0
    log("\n" + "=" * 70)
    log(f"✅ DONE (MAP) | N={len(ok)} | Acc={acc*100:.2f}% | NO_DETECT={nd*100:.2f}%")
    log("=" * 70)
    return rows
