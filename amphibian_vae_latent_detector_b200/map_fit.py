"""The Gaussian-MAP detector fit (08b_fit_map_detector.py:255-319) as host logic over device ops (row N1).

``ops`` supplies ``centroid_accumulate(Z, label, K) -> (sum f64 [K,D], cnt i64 [K])``,
``cov_accumulate(Z, label, mean f32 [K,D], k_sel) -> f64 [D,D]`` and
``map_score(Z, means, precision, a_const, log_prior, tau) -> (pred, best, scores)``.  The O(N D^2) work (means, second
moments, scoring every latent) runs on the GPU; the O(D^3) algebra on the tiny D x D matrices (shrinkage, eps I,
slogdet, inverse) is the reference's own numpy sequence in the reference's own dtypes, on the host.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch


@dataclass
class MapFit:
    """What 08b writes under ``map_detector`` (08b:322-351)."""
    species: List[str]               # sorted names of the species that have latents
    index: np.ndarray                # their label indices
    means: np.ndarray                # [K, D] float32
    cov: np.ndarray                  # [K, D, D] float32
    precision: np.ndarray            # [K, D, D] float32
    logdet_cov: np.ndarray           # [K] float64
    priors: np.ndarray               # [K] float64
    tau: Optional[float]
    counts: np.ndarray               # [K] int64
    scores_true: Optional[np.ndarray] = None   # float64, score of every latent under its own class (tau calibration)

    def constants(self):
        D = self.means.shape[1]
        a = self.logdet_cov + D * float(np.log(2.0 * np.pi))
        lp = np.log(self.priors + 1e-12)
        return a.astype(np.float64), lp.astype(np.float64)


def inv_and_logdet(cov: np.ndarray):
    """map_detector_core.py:306-316 (host, D x D)."""
    sign, ld = np.linalg.slogdet(cov)
    if sign <= 0:
        d = cov.shape[0]
        cov2 = cov + (1e-3 * np.eye(d, dtype=cov.dtype))
        sign, ld = np.linalg.slogdet(cov2)
        if sign <= 0:
            raise RuntimeError("Covarianza no PD incluso tras regularización.")
        cov = cov2
    return np.linalg.inv(cov).astype(np.float32), float(ld)


def regularise_cov(cov: np.ndarray, eps: float, shrink: float, cov_structure: str) -> np.ndarray:
    """The part of ``estimate_cov`` (08b:60-81) after ``np.cov``: diag, shrinkage, eps I, in float32."""
    d = cov.shape[0]
    cov = cov.astype(np.float32)
    if cov_structure == "diag":
        cov = np.diag(np.diag(cov)).astype(np.float32)
    if shrink > 0:
        avg_var = float(np.mean(np.diag(cov))) if d > 0 else 1.0
        cov = (1.0 - shrink) * cov + shrink * (avg_var * np.eye(d, dtype=np.float32))
    cov = cov + (eps * np.eye(d, dtype=np.float32))
    return cov.astype(np.float32)


def _np_cov(S: np.ndarray, s1: np.ndarray, n: int) -> np.ndarray:
    """``np.cov(Zc, rowvar=False, bias=False)`` from float64 sums: S = sum zc zc^T, s1 = sum zc (np.cov re-centres)."""
    d = S.shape[0]
    if n < 2:
        return np.eye(d, dtype=np.float32)
    m = s1 / n
    return ((S - n * np.outer(m, m)) / (n - 1)).astype(np.float32)


def fit_map(ops, Z: torch.Tensor, label: torch.Tensor, species_names: Sequence[str], *, cov_type: str = "lda",
            cov_structure: str = "full", priors: str = "empirical", eps: float = 1e-6, shrink: float = 0.0,
            set_tau_q: Optional[float] = None, group=None) -> MapFit:
    import torch.distributed as dist

    if not (0.0 <= shrink <= 1.0):
        raise ValueError("shrink debe estar en [0,1].")                                    # 08b:131-132
    if set_tau_q is not None and not (0.0 < float(set_tau_q) < 1.0):
        raise ValueError("set_tau_q debe estar en (0,1).")                                 # 08b:133-134
    K_all, D = len(species_names), Z.shape[1]
    distributed = group is not None and dist.get_world_size(group) > 1
    sums, cnts = ops.centroid_accumulate(Z, label, K_all)
    if distributed:
        packed = torch.cat([sums.reshape(-1), cnts.to(torch.float64)])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        sums, cnts = packed[:K_all * D].reshape(K_all, D), packed[K_all * D:].round().to(torch.int64)
    counts_all = cnts.cpu().numpy()
    present = [i for i in np.argsort(np.array(species_names, dtype=object), kind="stable") if counts_all[i] > 0]
    if not present:
        raise RuntimeError("No se codificó ninguna especie.")
    means_all = (sums / cnts.clamp_min(1).to(torch.float64)[:, None]).to(torch.float32)      # np.mean(...).astype(f32)
    sums_np, means_np = sums.cpu().numpy(), means_all.cpu().numpy()
    resid = sums_np - counts_all[:, None] * means_np.astype(np.float64)                      # sum of centred rows per class

    def second_moment(k_sel: int) -> np.ndarray:
        S = ops.cov_accumulate(Z, label, means_all, k_sel)
        if distributed:
            dist.all_reduce(S, op=dist.ReduceOp.SUM, group=group)
        return S.cpu().numpy()

    K = len(present)
    cov = np.zeros((K, D, D), dtype=np.float32)
    prec = np.zeros((K, D, D), dtype=np.float32)
    logdet = np.zeros(K, dtype=np.float64)
    if cov_type == "lda":
        n_tot = int(counts_all[present].sum())
        S = second_moment(-1)
        c = regularise_cov(_np_cov(S, resid[present].sum(axis=0), n_tot), float(eps), float(shrink), cov_structure)
        p, ld = inv_and_logdet(c)
        cov[:], prec[:], logdet[:] = c, p, ld
    elif cov_type == "qda":
        for j, i in enumerate(present):
            c = regularise_cov(_np_cov(second_moment(int(i)), resid[i], int(counts_all[i])), float(eps), float(shrink),
                               cov_structure)
            prec[j], logdet[j] = inv_and_logdet(c)
            cov[j] = c
    else:
        raise ValueError(cov_type)
    if priors == "uniform":
        pri = np.full(K, 1.0 / K)
    else:
        pri = counts_all[present].astype(np.float64) / float(counts_all[present].sum())
    fit = MapFit([species_names[i] for i in present], np.array(present, dtype=np.int32), means_np[present], cov, prec,
                 logdet, pri, None, counts_all[present])
    if set_tau_q is not None:
        # score of every latent under its own class (08b:298-319), then the quantile of all of them
        _, _, scores = ops.map_score(Z, fit, want_scores=True)
        remap = torch.full((K_all,), -1, dtype=torch.int64, device=label.device)
        remap[torch.as_tensor(present, device=label.device)] = torch.arange(K, device=label.device)
        col = remap[label.clamp_min(0).long()]
        ok = (label >= 0) & (col >= 0)
        true_scores = scores[ok].gather(1, col[ok][:, None])[:, 0]
        if distributed:
            n_local = torch.tensor([true_scores.shape[0]], dtype=torch.int64, device=true_scores.device)
            sizes = [torch.zeros_like(n_local) for _ in range(dist.get_world_size(group))]
            dist.all_gather(sizes, n_local, group=group)
            n_max = max(int(s.item()) for s in sizes)
            pad = torch.full((n_max,), float("nan"), dtype=torch.float64, device=true_scores.device)
            pad[:true_scores.shape[0]] = true_scores
            allv = torch.empty(n_max * len(sizes), dtype=torch.float64, device=true_scores.device)
            dist.all_gather_into_tensor(allv, pad, group=group)
            true_scores = allv[~torch.isnan(allv)]
        fit.scores_true = true_scores.cpu().numpy()
        fit.tau = float(np.quantile(fit.scores_true, float(set_tau_q)))
    return fit
