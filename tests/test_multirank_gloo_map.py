"""CPU, world size 2, gloo: the multi-rank host logic of the Gaussian-MAP fit (map_fit.fit_map, row N1 x row (e)) --
one packed all-reduce of per-species sums / counts, an all-reduce of the float64 second moments, an all-gather of the ragged
true-class scores for the tau quantile.  Every rank must end with bit-identical parameters, equal to the single-rank fit and
to the fixtures the reference's own 08b functions produced (tests/golden/map.npz).  Device ops = the numpy double of
tests/test_map_cli_host.py (defined in tests/, never importable by the product)."""
import json
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parents[1]
for p in (REPO, REPO / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

SPECIES = ["Batrachyla_leptopus", "Batrachyla_taeniata", "Calyptocephalella_gayi", "Pleurodema_thaul"]
CASES = ("lda_full", "qda_full")
KEYS = ("means", "cov", "precision", "logdet_cov", "priors", "counts", "tau", "scores_true_sorted")


def _case(tag):
    from oracle.make_golden_map import latents
    c = json.loads((REPO / "tests" / "golden" / "map_meta.json").read_text())["cases"][tag]
    Z, lab = latents(c["n"], c["d"], c["seed"])
    return c, Z, lab


def _fit(tag, Z, lab, group=None):
    from amphibian_vae_latent_detector_b200.map_fit import fit_map
    from test_map_cli_host import OracleEngine
    c, _, _ = _case(tag)
    fit = fit_map(OracleEngine(), torch.from_numpy(Z), torch.from_numpy(lab), SPECIES, cov_type=c["cov_type"],
                  cov_structure=c["cov_structure"], eps=c["eps"], shrink=c["shrink"], set_tau_q=c["tau_q"], group=group)
    return {"means": fit.means, "cov": fit.cov, "precision": fit.precision, "logdet_cov": fit.logdet_cov, "priors": fit.priors,
            "counts": fit.counts, "tau": np.array(fit.tau), "scores_true_sorted": np.sort(fit.scores_true)}


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    for tag in CASES:
        _, Z, lab = _case(tag)
        cut = int(0.37 * Z.shape[0])                      # ragged shards
        sl = slice(0, cut) if rank == 0 else slice(cut, None)
        np.savez(Path(out_dir) / f"{tag}_rank{rank}.npz", **_fit(tag, Z[sl], lab[sl], group=dist.group.WORLD))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_map_fit_equals_single_rank_and_reference(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    g = np.load(REPO / "tests" / "golden" / "map.npz")
    for tag in CASES:
        r0, r1 = np.load(tmp_path / f"{tag}_rank0.npz"), np.load(tmp_path / f"{tag}_rank1.npz")
        for key in KEYS:
            assert np.array_equal(r0[key], r1[key]), (tag, key)                 # bit-identical on all ranks
        c, Z, lab = _case(tag)
        single = _fit(tag, Z, lab)
        assert np.array_equal(single["counts"], r0["counts"])
        assert np.allclose(single["means"], r0["means"], rtol=1e-6, atol=1e-7)
        scale = np.abs(single["cov"]).max()
        assert np.max(np.abs(single["cov"] - r0["cov"])) <= 1e-6 * scale        # float64 sums in a different order
        assert np.allclose(single["logdet_cov"], r0["logdet_cov"], rtol=1e-6, atol=1e-5)
        assert float(r0["tau"]) == pytest.approx(float(single["tau"]), rel=1e-5)
        assert r0["scores_true_sorted"].shape == (c["n"],)                      # every latent of both shards was gathered
        # the reference's own estimate_cov / inv_and_logdet / scores on the same latents
        assert np.max(np.abs(r0["cov"] - g[f"{tag}_cov"])) <= 1e-5 * scale
        assert np.allclose(r0["logdet_cov"], g[f"{tag}_logdet"], rtol=1e-4, atol=1e-3)
        assert np.allclose(r0["means"], g[f"{tag}_means"], rtol=2e-5, atol=1e-5)
        assert float(r0["tau"]) == pytest.approx(c["tau"], rel=1e-3)
