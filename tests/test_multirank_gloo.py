"""CPU, world size 2, gloo: the multi-rank host logic of the radial fit (radial_fit.fit_radial) -- one packed
all-reduce of per-species sums/counts, all-gather of ragged radii blocks, identical thresholds on every rank equal
to the single-rank result and to the oracle.  The device ops are replaced by a numpy test double (defined HERE,
in tests/, never importable by the product) so the collectives can run without a GPU."""
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
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


class NumpyOps:
    """Test double for Engine's centroid_accumulate / radii / order_stats (CPU tensors)."""

    def centroid_accumulate(self, Z, label, K):
        Zn, ln = Z.numpy().astype(np.float64), label.numpy()
        s = np.zeros((K, Zn.shape[1]))
        c = np.zeros(K, dtype=np.int64)
        for k in range(K):
            s[k] = Zn[ln == k].sum(axis=0)
            c[k] = int((ln == k).sum())
        return torch.from_numpy(s), torch.from_numpy(c)

    def radii(self, Z, cent):
        d = Z[:, None, :] - cent[None, :, :]
        return torch.sqrt((d * d).sum(dim=2)).to(torch.float32)

    def order_stats(self, radii, label, queries):
        r, l = radii.numpy(), label.numpy()
        out = np.empty(len(queries), dtype=np.float32)
        for i, (k, side, rank) in enumerate(queries):
            sel = (l == k) if side == 0 else ((l != k) & (l >= 0))
            out[i] = np.sort(r[sel, k])[rank]
        return out


def _data(n=3001, d=32, seed=4):
    rng = np.random.default_rng(seed)
    cents = 3.0 * rng.standard_normal((4, d))
    labels = rng.integers(0, 4, n).astype(np.int32)
    Z = (cents[labels] + rng.standard_normal((n, d))).astype(np.float32)
    return Z, labels


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from amphibian_vae_latent_detector_b200.radial_fit import fit_radial
    Z, labels = _data()
    cut = 1300                                         # ragged shards: 1300 and 1701 rows
    sl = slice(0, cut) if rank == 0 else slice(cut, None)
    fit = fit_radial(NumpyOps(), torch.from_numpy(Z[sl]), torch.from_numpy(labels[sl]), 4, 0.95,
                     (0.10, 0.15, 0.20, 0.25), group=dist.group.WORLD)
    # the same with a block size all ranks agree on beforehand (no size exchange, no host sync before the gather)
    fit2 = fit_radial(NumpyOps(), torch.from_numpy(Z[sl]), torch.from_numpy(labels[sl]), 4, 0.95,
                      (0.10, 0.15, 0.20, 0.25), group=dist.group.WORLD, shard_rows=1701)
    assert np.array_equal(fit.rk, fit2.rk, equal_nan=True) and np.array_equal(fit.centroids, fit2.centroids, equal_nan=True)
    assert np.array_equal(fit.summaries["out"], fit2.summaries["out"], equal_nan=True)
    np.savez(Path(out_dir) / f"rank{rank}.npz", centroids=fit.centroids, rk=fit.rk, rk_in=fit.rk_in, rk_out=fit.rk_out,
             counts=fit.counts, s_in=fit.summaries["in"], s_out=fit.summaries["out"])
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_fit_equals_single_rank_and_oracle(tmp_path):
    from amphibian_vae_latent_detector_b200.radial_fit import fit_radial
    from oracle import hotpath as hp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for key in r0.files:
        assert np.array_equal(r0[key], r1[key], equal_nan=True), key          # bit-identical on all ranks
    Z, labels = _data()
    single = fit_radial(NumpyOps(), torch.from_numpy(Z), torch.from_numpy(labels), 4, 0.95, (0.10, 0.15, 0.20, 0.25))
    assert np.array_equal(single.counts, r0["counts"])
    assert np.allclose(single.centroids, r0["centroids"], rtol=1e-6, atol=1e-7)
    assert np.allclose(single.rk, r0["rk"], rtol=1e-6)                       # 1e-6 relative across rank counts
    for qi, q in enumerate((0.10, 0.15, 0.20, 0.25)):
        cent, rk, rk_in, rk_out = hp.fit_radial(Z, labels, 4, 0.95, q)
        assert np.allclose(r0["rk"][qi], rk, rtol=1e-5)
        assert np.allclose(r0["rk_in"], rk_in, rtol=1e-5) and np.allclose(r0["rk_out"][qi], rk_out, rtol=1e-5)
        assert np.allclose(r0["centroids"], cent, rtol=1e-4, atol=1e-5)


def test_single_rank_matches_reference_fixture(golden_meta):
    """host logic + numpy double vs the reference-made fit fixtures (exact order statistics + numpy lerp => equal floats)."""
    from amphibian_vae_latent_detector_b200.radial_fit import fit_radial
    rng = np.random.default_rng(31)
    cents = 3.0 * rng.standard_normal((4, 128))
    labels = (np.arange(400) % 4).astype(np.int32)
    Z = (cents[labels] + rng.standard_normal((400, 128))).astype(np.float32)
    fit = fit_radial(NumpyOps(), torch.from_numpy(Z), torch.from_numpy(labels), 4, 0.95, (0.01, 0.10, 0.25))
    for qi, q in enumerate((0.01, 0.10, 0.25)):
        for k in range(4):
            g = golden_meta["fit"][f"small_q{q:.2f}_k{k}"]
            assert fit.rk[qi, k] == pytest.approx(g["rk"], rel=2e-6)
            assert fit.rk_in[k] == pytest.approx(g["rk_in"], rel=2e-6)
