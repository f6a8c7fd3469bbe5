"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's hot path (array in, array out).

The reference works file-at-a-time (WAV path in, latent out).  These functions restate the
same arithmetic on in-memory arrays so the CUDA path can be checked on the GPU box, where
``/root/reference`` does not exist.  Each function cites the reference lines it follows
(paths relative to ``/root/reference/latent_space_exploration``).  They are pinned against the
reference's own code (run in the build container through ``oracle/ref_import.py``) by the
fixtures in ``tests/golden/`` -- see ``oracle/make_golden.py`` and
``tests/test_oracle_vs_golden.py``.

numpy-version note: the in-container oracle executes numpy 2.3.5 scalar-promotion rules
(``rms + eps`` and ``0.05 / (...)`` stay float32); the reference pins numpy 1.26.4 which does
those two scalar ops in float64 (1-ulp scale difference in ~37 % of chunks, SURVEY.md
section 7 hard part 1).  ``rms_normalize(..., numpy1_scalars=True)`` restates the 1.26 variant.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import librosa_port as lp

PRIORITY_ORDER: List[str] = [  # 09_evaluate_wav_detection.py:61-66
    "Batrachyla_leptopus",
    "Batrachyla_taeniata",
    "Calyptocephalella_gayi",
    "Pleurodema_thaul",
]


# ----------------------------------------------------------------------------------------
# R1: 00_normalize_dataset_rms.py:29-38
# ----------------------------------------------------------------------------------------
def rms_normalize(y: np.ndarray, target_rms=0.05, rms_min=1e-4, eps=1e-8,
                  numpy1_scalars: bool = False) -> Tuple[np.ndarray, bool]:
    rms = np.sqrt(np.mean(y ** 2))
    if (float(rms) < rms_min) if numpy1_scalars else (rms < rms_min):      # numpy 1.x compares in float64 as well
        return y, False
    if numpy1_scalars:  # numpy 1.26.4 value-based casting: float32 scalar (op) python float -> float64
        scale = np.float32(target_rms / (float(rms) + eps))
        y_norm = y * scale
    else:
        y_norm = y * (target_rms / (rms + eps))
    y_norm = np.clip(y_norm, -1.0, 1.0)
    return y_norm, True


def rms_normalize_batch(x: np.ndarray, target_rms=0.05, rms_min=1e-4, eps=1e-8,
                        pcm16: bool = False, numpy1_scalars: bool = False) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Row-wise R1 (+ optional R2 ``sf.write``/``librosa.load`` PCM_16 round trip,
    00_normalize_dataset_rms.py:55-57 -> map_detector_core.py:210).  Returns ``(y, ok, rms)``."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    ok = np.zeros(x.shape[0], dtype=np.uint8)
    rms = np.zeros(x.shape[0], dtype=np.float32)
    for i in range(x.shape[0]):
        row = x[i]
        rms[i] = np.sqrt(np.mean(row ** 2))
        yn, good = rms_normalize(row, target_rms, rms_min, eps, numpy1_scalars=numpy1_scalars)
        ok[i] = 1 if good else 0
        y[i] = lp.pcm16_roundtrip(yn) if pcm16 else yn
    return y, ok, rms


# ----------------------------------------------------------------------------------------
# M1-M5: map_detector_core.py:185-237 (= 07:206-257, 08:201-247, 09:229-282)
# ----------------------------------------------------------------------------------------
def crop_or_pad_time(mel: np.ndarray, target_frames: int) -> np.ndarray:
    """map_detector_core.py:185-195."""
    _, T = mel.shape
    if T == target_frames:
        return mel
    if T > target_frames:
        start = (T - target_frames) // 2
        return mel[:, start:start + target_frames]
    pad_total = target_frames - T
    pad_left = pad_total // 2
    return np.pad(mel, ((0, 0), (pad_left, pad_total - pad_left)), mode="constant")


def fix_length(y: np.ndarray, sr: int, duration: float) -> np.ndarray:
    """map_detector_core.py:212-217."""
    if duration > 0:
        target_len = int(sr * duration)
        if y.shape[0] < target_len:
            y = np.pad(y, (0, target_len - y.shape[0]), mode="constant")
        else:
            y = y[:target_len]
    return y


def mel_power(y: np.ndarray, *, sr=48000, n_mels=64, fmin=150.0, fmax=15000.0, hop_length=384,
              n_fft=2048) -> np.ndarray:
    """map_detector_core.py:219-228 -> float32 ``[n_mels, F]``."""
    return lp.melspectrogram(y=y, sr=sr, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels,
                             fmin=fmin, fmax=fmax, power=2.0)


def logmel_features(y: np.ndarray, *, sr=48000, duration=0.0, n_mels=64, fmin=150.0, fmax=15000.0,
                    hop_length=384, n_fft=2048, target_frames=192) -> np.ndarray:
    """``wav_to_mel`` after the file load (map_detector_core.py:212-237) -> float32 ``[n_mels, T]``."""
    y = fix_length(np.asarray(y, dtype=np.float32), sr, duration)
    S = mel_power(y, sr=sr, n_mels=n_mels, fmin=fmin, fmax=fmax, hop_length=hop_length, n_fft=n_fft)
    S_db = lp.power_to_db(S, ref=np.max)                       # :229
    S_db = (S_db - S_db.mean()) / (S_db.std() + 1e-8)          # :231-232
    S_db = crop_or_pad_time(S_db, target_frames=target_frames)  # :235
    return np.ascontiguousarray(S_db, dtype=np.float32)


def logmel_features_batch(y: np.ndarray, **kw) -> np.ndarray:
    """Rows of ``y`` -> encoder input layout ``[n, T, M]`` (map_detector_core.py:267-268: ``mel.T``)."""
    return np.stack([logmel_features(row, **kw).T for row in y], axis=0)


# ----------------------------------------------------------------------------------------
# E0-E2: map_detector_core.py:240-300
# ----------------------------------------------------------------------------------------
def extract_latent(out: Any):
    """Output -> ``[B, D]`` tensor, map_detector_core.py:272-294 (07:264-297 adds key 'enc')."""
    import torch

    if isinstance(out, torch.Tensor):
        t = out
    elif isinstance(out, (list, tuple)):
        t = next((z for z in out if isinstance(z, torch.Tensor)), None)
        if t is None:
            raise RuntimeError("encoder output is a tuple/list without tensors")
    elif isinstance(out, dict):
        t = None
        for k in ("z", "latent", "mu", "mean", "embedding"):
            if k in out and isinstance(out[k], torch.Tensor):
                t = out[k]
                break
        if t is None:
            t = next((v for v in out.values() if isinstance(v, torch.Tensor)), None)
        if t is None:
            raise RuntimeError("encoder output is a dict without tensors")
    else:
        raise RuntimeError(f"cannot interpret encoder output: {type(out)}")
    if t.ndim == 3:
        t = t.mean(dim=1)
    if t.ndim > 2:
        t = t.view(t.shape[0], -1)
    return t


def encode_features(encoder, feat_mt: np.ndarray) -> np.ndarray:
    """One chunk, batch 1, exactly as the reference runs it (map_detector_core.py:267-300)."""
    import torch

    with torch.no_grad():
        mel = torch.tensor(feat_mt, dtype=torch.float32)       # [M, T]
        x = mel.T.unsqueeze(0).unsqueeze(0)                    # [1, 1, T, M]
        z = extract_latent(encoder(x)).detach().cpu().numpy()
    if z.shape[0] != 1:
        raise RuntimeError(f"expected batch=1, got {z.shape}")
    return z[0].astype(np.float32)


def encode_array_to_latent(encoder, y: np.ndarray, **mel_kw) -> np.ndarray:
    return encode_features(encoder, logmel_features(y, **mel_kw))


def encode_batch(encoder, y: np.ndarray, **mel_kw) -> np.ndarray:
    """Per-file loop, batch 1: the reference's execution model (08:488-506)."""
    return np.stack([encode_array_to_latent(encoder, row, **mel_kw) for row in y], axis=0)


# ----------------------------------------------------------------------------------------
# F1-F3: 08_fit_radial_detector.py:105-123, :310-333, :530-558
# ----------------------------------------------------------------------------------------
def l2_norm_rows(x: np.ndarray) -> np.ndarray:
    return np.sqrt(np.sum(x * x, axis=1))


def quantile_safe(x: np.ndarray, q: float) -> float:
    if x.size == 0:
        return 0.0
    return float(np.quantile(x, q))


def summarize_dist(x: np.ndarray) -> Dict[str, float]:
    if x.size == 0:
        return {"min": float("nan"), "p50": float("nan"), "p90": float("nan"), "max": float("nan")}
    return {"min": float(np.min(x)), "p50": float(np.quantile(x, 0.50)),
            "p90": float(np.quantile(x, 0.90)), "max": float(np.max(x))}


def fit_species_with_fp_control(Z_in: np.ndarray, Z_out: Optional[np.ndarray], q_in: float, q_out: float):
    mu = np.mean(Z_in, axis=0).astype(np.float32)
    rho_in = l2_norm_rows(Z_in - mu[None, :])
    rk_in = quantile_safe(rho_in, q_in)
    if Z_out is None or Z_out.size == 0:
        rho_out = np.array([], dtype=np.float32)
        rk_out = float("inf")
    else:
        rho_out = l2_norm_rows(Z_out - mu[None, :])
        rk_out = quantile_safe(rho_out, q_out)
    rk = float(min(rk_in, rk_out))
    extra = {"rho_in_summary": summarize_dist(rho_in), "rho_out_summary": summarize_dist(rho_out)}
    return mu, rk, rk_in, rk_out, extra


def fit_radial(Z: np.ndarray, labels: np.ndarray, K: int, q_in: float, q_out: float):
    """The fit loop 08:530-558 on a labelled latent matrix.  Returns ``(centroids[K,D] f32,
    rk[K] f64, rk_in[K] f64, rk_out[K] f64)``; species without members get NaN centroid / rk."""
    D = Z.shape[1]
    cent = np.full((K, D), np.nan, dtype=np.float32)
    rk = np.full(K, np.nan)
    rk_in = np.full(K, np.nan)
    rk_out = np.full(K, np.nan)
    for k in range(K):
        Z_in = Z[labels == k]
        if Z_in.shape[0] == 0:
            continue
        Z_out = Z[(labels != k) & (labels >= 0)]
        mu, r, ri, ro, _ = fit_species_with_fp_control(Z_in, Z_out if Z_out.shape[0] else None, q_in, q_out)
        cent[k], rk[k], rk_in[k], rk_out[k] = mu, r, ri, ro
    return cent, rk, rk_in, rk_out


# ----------------------------------------------------------------------------------------
# D2: 09_evaluate_wav_detection.py:354-355, :416-436; 10_benchmark_folder_detection.py:175-199
# ----------------------------------------------------------------------------------------
def l2(a: np.ndarray) -> float:
    return float(np.sqrt(np.sum(a * a)))


def decide_one(z: np.ndarray, centroids: Dict[str, np.ndarray], thresholds: Dict[str, float],
               priority: Sequence[str] = PRIORITY_ORDER) -> Tuple[bool, Optional[str], float]:
    accepted: List[str] = []
    best_d = float("inf")
    for sp, mu in centroids.items():
        if sp not in thresholds:
            continue
        rk = float(thresholds[sp])
        if mu.shape[0] != z.shape[0]:
            continue
        d = float(l2(z - mu))
        best_d = min(best_d, d)
        if d <= rk:
            accepted.append(sp)
    if not accepted:
        return False, None, best_d
    for sp in priority:
        if sp in accepted:
            return True, sp, best_d
    return True, sorted(accepted)[0], best_d


def decide_batch(Z: np.ndarray, species: Sequence[str], centroids: np.ndarray, thresholds: np.ndarray,
                 priority: Sequence[str] = PRIORITY_ORDER) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Per-row D2 -> ``(pred[N] int32 index into species or -1, best_d[N] f32, radii[N,K] f32)``."""
    cd = {sp: centroids[i] for i, sp in enumerate(species)}
    td = {sp: float(thresholds[i]) for i, sp in enumerate(species)}
    pred = np.full(Z.shape[0], -1, dtype=np.int32)
    best = np.zeros(Z.shape[0], dtype=np.float32)
    radii = np.zeros((Z.shape[0], len(species)), dtype=np.float32)
    for i, z in enumerate(Z):
        det, sp, bd = decide_one(z, cd, td, priority)
        pred[i] = species.index(sp) if det else -1
        best[i] = bd
        for k, s in enumerate(species):
            radii[i, k] = l2(z - cd[s])
    return pred, best, radii


# ----------------------------------------------------------------------------------------
# N1: Gaussian-MAP detector
#   map_detector_core.py:306-323 (inv_and_logdet, gaussian_logpdf_from_precision)
#   08b_fit_map_detector.py:60-81 (estimate_cov), :255-319 (fit)
#   09n_evaluate_wav_detection.py:114-140 / 10b_benchmark_folder_detection_map.py:146-169 (decision)
# ----------------------------------------------------------------------------------------
def inv_and_logdet(cov: np.ndarray) -> Tuple[np.ndarray, float]:
    """core:306-316."""
    sign, ld = np.linalg.slogdet(cov)
    if sign <= 0:
        d = cov.shape[0]
        cov2 = cov + (1e-3 * np.eye(d, dtype=cov.dtype))
        sign, ld = np.linalg.slogdet(cov2)
        if sign <= 0:
            raise RuntimeError("Covarianza no PD incluso tras regularización.")
        cov = cov2
    prec = np.linalg.inv(cov).astype(np.float32)
    return prec, float(ld)


def gaussian_logpdf_from_precision(z: np.ndarray, mu: np.ndarray, prec: np.ndarray, logdet_cov: float) -> float:
    """core:319-323."""
    d = int(z.shape[0])
    diff = (z - mu).astype(np.float32)
    quad = float(diff.T @ prec @ diff)
    return -0.5 * (quad + float(logdet_cov) + d * float(np.log(2.0 * np.pi)))


def estimate_cov(Z: np.ndarray, eps: float, shrink: float, cov_structure: str) -> np.ndarray:
    """08b:60-81."""
    n, d = Z.shape
    if n < 2:
        cov = np.eye(d, dtype=np.float32)
    else:
        cov = np.cov(Z, rowvar=False, bias=False).astype(np.float32)
    if cov_structure == "diag":
        cov = np.diag(np.diag(cov)).astype(np.float32)
    if shrink > 0:
        avg_var = float(np.mean(np.diag(cov))) if d > 0 else 1.0
        cov = (1.0 - shrink) * cov + shrink * (avg_var * np.eye(d, dtype=np.float32))
    cov = cov + (eps * np.eye(d, dtype=np.float32))
    return cov.astype(np.float32)


def fit_map(Z_by_species: Dict[str, np.ndarray], *, cov_type="lda", cov_structure="full", priors="empirical",
            eps=1e-6, shrink=0.0, set_tau_q=None):
    """08b:255-319 on in-memory latents -> dict(means, cov, precision, logdet_cov, priors, tau, scores_true)."""
    species = sorted(Z_by_species.keys())
    K = len(species)
    if priors == "uniform":
        pri = {sp: 1.0 / K for sp in species}
    else:
        total = float(sum(Z_by_species[sp].shape[0] for sp in species))
        pri = {sp: float(Z_by_species[sp].shape[0]) / total for sp in species}
    means = {sp: np.mean(Z_by_species[sp], axis=0).astype(np.float32) for sp in species}
    covs, precs, logdets = {}, {}, {}
    if cov_type == "lda":
        Zc = np.concatenate([Z_by_species[sp] - means[sp][None, :] for sp in species], axis=0)
        cov_shared = estimate_cov(Zc, eps=float(eps), shrink=float(shrink), cov_structure=cov_structure)
        prec_shared, logdet_shared = inv_and_logdet(cov_shared)
        for sp in species:
            covs[sp], precs[sp], logdets[sp] = cov_shared, prec_shared, logdet_shared
    else:
        for sp in species:
            Zc = Z_by_species[sp] - means[sp][None, :]
            covs[sp] = estimate_cov(Zc, eps=float(eps), shrink=float(shrink), cov_structure=cov_structure)
            precs[sp], logdets[sp] = inv_and_logdet(covs[sp])
    scores_true: List[float] = []
    for sp in species:
        lp = float(np.log(pri[sp] + 1e-12))
        scores_true.extend(gaussian_logpdf_from_precision(z, means[sp], precs[sp], logdets[sp]) + lp
                           for z in Z_by_species[sp])
    arr = np.array(scores_true, dtype=np.float64)
    tau = float(np.quantile(arr, float(set_tau_q))) if set_tau_q is not None else None
    return dict(species=species, means=means, cov=covs, precision=precs, logdet_cov=logdets, priors=pri, tau=tau,
                scores_true=arr)


def decide_map_one(z: np.ndarray, species: Sequence[str], means, precisions, logdets, priors, tau):
    """09n:114-140 -> (detected, species | None, best_score)."""
    best_sp, best_score = None, -float("inf")
    for sp in species:
        mu, prec = means[sp], precisions[sp]
        if mu.shape[0] != z.shape[0] or prec.shape[0] != z.shape[0] or prec.shape[1] != z.shape[0]:
            continue
        lp = float(np.log(float(priors.get(sp, 1e-12)) + 1e-12))
        s = gaussian_logpdf_from_precision(z, mu, prec, logdets[sp]) + lp
        if s > best_score:
            best_score, best_sp = s, sp
    if best_sp is None:
        return False, None, best_score
    if tau is not None and best_score < float(tau):
        return False, None, best_score
    return True, best_sp, best_score
