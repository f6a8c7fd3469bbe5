"""The radial fit (08_fit_radial_detector.py:310-333, :530-558) as host logic over four device ops.

``ops`` supplies ``centroid_accumulate(Z, label, K) -> (sum f64 [K,D], cnt i64 [K])``,
``radii(Z, centroid) -> [n,K] f32`` and ``order_stats(radii, label, queries) -> np.float32[n_q]``.
In the product ``ops`` is :class:`engine.Engine` (CUDA kernels behind ``libavld.so``); the host logic
here -- collectives, which ranks are needed, numpy's linear interpolation -- is device independent so
that the multi-rank path can be exercised on CPU with ``gloo`` in ``tests/``.

One pass computes the radii of every latent to every centroid once and then serves *every* quantile of
a q_out grid (run_qout_grid.sh:13 refits per grid point; the radii do not depend on q).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import quantile as _q


@dataclass
class RadialFit:
    """What 08 writes under ``radial_detector`` (08:561-583), for every q_out of the grid."""
    centroids: np.ndarray          # [K, D] float32 (NaN rows for species without members)
    counts: np.ndarray             # [K] int64
    rk_in: np.ndarray              # [K] float64
    rk_out: np.ndarray             # [n_qout, K] float64 (inf where there are no out-of-class rows, 08:321-323)
    rk: np.ndarray                 # [n_qout, K] float64 = min(rk_in, rk_out) (08:328)
    q_in: float
    q_out: Tuple[float, ...]
    summaries: Dict[str, np.ndarray]   # 'in' / 'out' -> [K, 4] = (min, p50, p90, max), summarize_dist 08:115-123
    radii_local: Optional[torch.Tensor] = None   # [n_local, K] radii of this rank's rows to the fitted centroids


def all_gather_rows(radii: torch.Tensor, label: torch.Tensor, group, shard_rows: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather of ragged row blocks: pad every rank to the largest block, mark the padding with
    label -1 (ignored by the selection kernels), gather radii and labels.  ``shard_rows`` = a block size every rank
    already agrees on (>= each rank's rows, e.g. ceil(n_total / world) for contiguous shards): no size exchange and no
    host synchronisation, the gather is stream-ordered behind the radii kernel."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if shard_rows is not None:
        if radii.shape[0] > shard_rows:
            raise ValueError(f"shard_rows = {shard_rows} is smaller than this rank's {radii.shape[0]} rows")
        n_max = int(shard_rows)
    else:
        n_local = torch.tensor([radii.shape[0]], dtype=torch.int64, device=radii.device)
        sizes = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(sizes, n_local, group=group)
        n_max = max(int(s.item()) for s in sizes)
    K = radii.shape[1]
    pr = torch.zeros(n_max, K, dtype=radii.dtype, device=radii.device)
    pl = torch.full((n_max,), -1, dtype=label.dtype, device=label.device)
    pr[:radii.shape[0]] = radii
    pl[:label.shape[0]] = label
    gr = torch.empty(world * n_max, K, dtype=radii.dtype, device=radii.device)
    gl = torch.empty(world * n_max, dtype=label.dtype, device=label.device)
    dist.all_gather_into_tensor(gr, pr, group=group)
    dist.all_gather_into_tensor(gl, pl, group=group)
    return gr, gl


def fit_radial(ops, Z: torch.Tensor, label: torch.Tensor, K: int, q_in: float = 0.95,
               q_out: float | Sequence[float] = 0.01, *, group=None, semantics: str = "numpy2",
               shard_rows: Optional[int] = None) -> RadialFit:
    """Centroids, in/out radius quantiles and thresholds for every species and every ``q_out``.

    With a ``torch.distributed`` ``group`` (one rank per GPU, rows of ``Z`` sharded): per-species
    sums/counts are all-reduced in ONE float64 message (so the centroid does not depend on the rank
    count), every rank forms the same centroids and computes its local radii, and radii + labels are
    all-gathered so that every rank selects the same exact order statistics: thresholds are
    bit-identical on all ranks and equal to the single-rank result.  With ``shard_rows`` (see :func:`all_gather_rows`) the
    device work -- centroid kernel, all-reduce, radii kernel, all-gather -- is enqueued without a host synchronisation in
    between; the host first looks at the counts when it plans the rank queries of the selection.
    """
    import torch.distributed as dist

    q_outs = tuple(float(q) for q in (q_out if isinstance(q_out, (list, tuple, np.ndarray)) else [q_out]))
    for q in (q_in, *q_outs):
        if not (0.0 < q < 1.0):
            raise ValueError("q_in and q_out must be in (0,1)")              # 08:369-372
    D = Z.shape[1]
    sums, cnts = ops.centroid_accumulate(Z, label, K)
    # group = "avld": the context's own NCCL communicator (Engine.comm_init -> avld_comm_* of the C ABI) instead of
    # torch.distributed -- the form a host without PyTorch's process groups uses; same exchanges, same results
    own_comm = isinstance(group, str) and group == "avld"
    if own_comm:
        if getattr(ops, "comm_world", 0) < 1:
            raise ValueError('group="avld" needs Engine.comm_init first')
        if shard_rows is None:
            raise ValueError('group="avld" needs shard_rows (a block size every rank agrees on)')
        ops.allreduce_centroids(sums, cnts)
    distributed = (not own_comm) and group is not None and dist.get_world_size(group) > 1
    if distributed:
        packed = torch.cat([sums.reshape(-1), cnts.to(torch.float64)])        # counts < 2^53 are exact in f64
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        sums = packed[:K * D].reshape(K, D)
        cnts = packed[K * D:].round().to(torch.int64)
    cent = (sums / cnts.clamp_min(1).to(torch.float64)[:, None]).to(torch.float32)   # mean(...).astype(f32), 08:316
    radii = ops.radii(Z, cent)
    radii_local = radii
    lab = label
    if distributed:
        radii, lab = all_gather_rows(radii, label, group, shard_rows)
    elif own_comm:
        radii, lab = ops.allgather_radii(radii, label, shard_rows)
    counts = cnts.cpu().numpy()            # first host read: everything above is already enqueued
    n_tot = int(counts.sum())

    queries: List[Tuple[int, int, int]] = []
    plan = []   # (kind, k, qi, index of prev, index of next, gamma)

    def add(k, side, n_side, q, kind, qi):
        prev, nxt, gamma = _q.neighbour_ranks(n_side, q, semantics)
        queries.extend([(k, side, prev), (k, side, nxt)])
        plan.append((kind, k, qi, len(queries) - 2, len(queries) - 1, gamma))

    for k in range(K):
        n_in = int(counts[k])
        n_out = n_tot - n_in
        if n_in > 0:
            add(k, 0, n_in, q_in, "in", 0)
            for j, q in enumerate((0.0, 0.5, 0.9, 1.0)):
                add(k, 0, n_in, q, "sin", j)
            if n_out > 0:
                for qi, q in enumerate(q_outs):
                    add(k, 1, n_out, q, "out", qi)
                for j, q in enumerate((0.0, 0.5, 0.9, 1.0)):
                    add(k, 1, n_out, q, "sout", j)
    vals = ops.order_stats(radii, lab, queries) if queries else np.zeros(0, np.float32)
    rk_in = np.full(K, np.nan)
    rk_out = np.full((len(q_outs), K), np.inf)
    summ = {"in": np.full((K, 4), np.nan), "out": np.full((K, 4), np.nan)}
    for kind, k, qi, a, b, gamma in plan:
        v = _q.lerp(float(vals[a]), float(vals[b]), gamma, semantics)
        if kind == "in":
            rk_in[k] = v
        elif kind == "out":
            rk_out[qi, k] = v
        elif kind == "sin":
            summ["in"][k, qi] = v
        else:
            summ["out"][k, qi] = v
    rk_out[:, np.isnan(rk_in)] = np.nan
    rk = np.minimum(rk_in[None, :], rk_out)
    cent_np = cent.cpu().numpy().copy()
    cent_np[counts == 0] = np.nan
    return RadialFit(cent_np, counts, rk_in, rk_out, rk, float(q_in), q_outs, summ, radii_local)
