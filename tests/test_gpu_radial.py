"""GPU: centroid sums, radii, exact order statistics, the fit (08:310-333, :530-558) and the decision
(09:416-436, 10:175-199) against the numpy oracle and the reference-made fixtures."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import hotpath as hp
from amphibian_vae_latent_detector_b200.engine import priority_ranks

pytestmark = pytest.mark.gpu


def _latents(n, d, seed, k=4):
    rng = np.random.default_rng(seed)
    cents = 3.0 * rng.standard_normal((k, d))
    labels = (np.arange(n) % k).astype(np.int32)
    Z = (cents[labels] + rng.standard_normal((n, d))).astype(np.float32)
    return Z, labels


def test_centroid_and_radii(engine3s):
    Z, lab = _latents(20000, 128, 33)
    lab[::97] = -1                                         # failed / unlabeled rows are skipped
    s, c = engine3s.centroid_accumulate(torch.from_numpy(Z).cuda(), torch.from_numpy(lab).cuda(), 4)
    for k in range(4):
        ref = Z[lab == k].astype(np.float64).sum(axis=0)
        assert int(c[k]) == int((lab == k).sum())
        assert np.allclose(s[k].cpu().numpy(), ref, rtol=1e-12, atol=1e-9)
    cent = np.stack([Z[lab == k].mean(axis=0) for k in range(4)]).astype(np.float32)
    r = engine3s.radii(torch.from_numpy(Z).cuda(), torch.from_numpy(cent).cuda()).cpu().numpy()
    ref = np.stack([hp.l2_norm_rows(Z - cent[k][None]) for k in range(4)], axis=1)
    assert np.max(np.abs(r - ref)) / np.max(ref) < 1e-6


@pytest.mark.parametrize("tag,n,d,seed", [("small", 400, 128, 31), ("tiny", 7, 16, 32), ("big", 20000, 128, 33)])
def test_fit_vs_reference_fixture(engine3s, golden_meta, tag, n, d, seed):
    """Same recipe as oracle/make_golden.py section D: thresholds vs the reference's fit_species_with_fp_control."""
    Z, lab = _latents(n, d, seed)
    grid = (0.01, 0.10, 0.15, 0.20, 0.25)
    fit = engine3s.fit_radial(torch.from_numpy(Z).cuda(), torch.from_numpy(lab).cuda(), 4, 0.95, grid)
    for qi, q in enumerate(grid):
        for k in range(4):
            g = golden_meta["fit"][f"{tag}_q{q:.2f}_k{k}"]
            assert fit.rk_in[k] == pytest.approx(g["rk_in"], rel=1e-5)
            assert fit.rk_out[qi, k] == pytest.approx(g["rk_out"], rel=1e-5)
            assert fit.rk[qi, k] == pytest.approx(g["rk"], rel=1e-5)
            assert np.allclose(fit.centroids[k][:4], g["mu_head"], rtol=1e-5, atol=1e-6)
            for j, key in enumerate(("min", "p50", "p90", "max")):
                assert fit.summaries["in"][k, j] == pytest.approx(g["extra"]["rho_in_summary"][key], rel=1e-5)
                assert fit.summaries["out"][k, j] == pytest.approx(g["extra"]["rho_out_summary"][key], rel=1e-5)


def test_order_stats_exact(engine3s):
    rng = np.random.default_rng(1)
    n, K = 100003, 4
    radii = np.abs(rng.standard_normal((n, K))).astype(np.float32) * 5
    radii[::11, 2] = radii[5, 2]                            # heavy ties
    lab = rng.integers(-1, K, n).astype(np.int32)
    queries = []
    for k in range(K):
        n_in, n_out = int((lab == k).sum()), int(((lab != k) & (lab >= 0)).sum())
        queries += [(k, 0, 0), (k, 0, n_in // 2), (k, 0, n_in - 1), (k, 1, 0), (k, 1, n_out // 3), (k, 1, n_out - 1)]
    got = engine3s.order_stats(torch.from_numpy(radii).cuda(), torch.from_numpy(lab).cuda(), queries)
    for (k, side, rank), v in zip(queries, got):
        col = radii[(lab == k) if side == 0 else ((lab != k) & (lab >= 0)), k]
        assert v == np.sort(col)[rank]


def test_empty_species_and_no_out(engine3s):
    Z, lab = _latents(1000, 32, 2)
    lab[lab == 2] = 0                                       # species 2 has no members
    fit = engine3s.fit_radial(torch.from_numpy(Z).cuda(), torch.from_numpy(lab).cuda(), 4, 0.95, 0.1)
    assert fit.counts[2] == 0 and np.isnan(fit.rk[0, 2]) and np.all(np.isnan(fit.centroids[2]))
    only = np.zeros(300, dtype=np.int32)                    # a single species: no out-of-class rows -> rk_out = inf
    fit1 = engine3s.fit_radial(torch.from_numpy(Z[:300]).cuda(), torch.from_numpy(only).cuda(), 1, 0.95, 0.1)
    assert np.isinf(fit1.rk_out[0, 0]) and fit1.rk[0, 0] == fit1.rk_in[0]
    ref = hp.fit_species_with_fp_control(Z[:300], None, 0.95, 0.1)
    assert fit1.rk[0, 0] == pytest.approx(ref[1], rel=1e-5)


def test_decide_vs_reference_fixture(engine3s, golden_meta):
    d = np.load(GOLDEN / "decision.npz")
    species = golden_meta["species"]
    Z = np.stack([np.load(GOLDEN / f"feat_{k}.npz")["z"] for k in golden_meta["decision_cases"]])
    prio = priority_ranks(species, hp.PRIORITY_ORDER)
    r = engine3s.radii(torch.from_numpy(Z).cuda(), torch.from_numpy(d["centroids"]).cuda())
    pred, best = engine3s.decide(r, torch.from_numpy(d["thresholds"]).cuda(), torch.from_numpy(prio).cuda())
    pred, best = pred.cpu().numpy(), best.cpu().numpy()
    for i, key in enumerate(golden_meta["decision_cases"]):
        g = golden_meta["decision"][key]
        assert (species[pred[i]] if pred[i] >= 0 else None) == g["species"], key
        assert best[i] == pytest.approx(g["best_d"], rel=1e-5)


def test_decide_random_vs_oracle(engine3s):
    species = ["Pleurodema_thaul", "zzz", "Batrachyla_taeniata", "aaa", "Batrachyla_leptopus"]   # shuffled priority
    Z, _ = _latents(5000, 64, 8, k=5)
    rng = np.random.default_rng(4)
    cent = (rng.standard_normal((5, 64)) * 3).astype(np.float32)
    thr = np.array([12.0, 40.0, 13.0, 40.0, np.nan])       # NaN threshold: never accepted, still counts for best_d (10:177-187)
    prio = priority_ranks(species, hp.PRIORITY_ORDER)
    r = engine3s.radii(torch.from_numpy(Z).cuda(), torch.from_numpy(cent).cuda())
    pred, best = engine3s.decide(r, torch.from_numpy(thr).cuda(), torch.from_numpy(prio).cuda())
    po, bo, ro = hp.decide_batch(Z, species, cent, thr)
    rr = r.cpu().numpy()
    near = np.any(np.abs(rr[:, :4] - thr[None, :4]) / thr[None, :4] <= 1e-5, axis=1)
    assert np.array_equal(pred.cpu().numpy()[~near], po[~near])
    assert np.allclose(best.cpu().numpy(), bo, rtol=1e-5)
    assert len(set(po.tolist())) >= 3


@pytest.mark.gpu
def test_own_nccl_collectives_match_single_gpu_fit():
    """avld_comm_init / avld_allreduce_centroids / avld_allgather_radii (NCCL through dlopen, no torch.distributed): the
    sharded fit on two GPUs is bit-identical to the single-GPU fit (tools/comm_check.py; needs two visible GPUs)."""
    import subprocess
    import sys
    from pathlib import Path
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    repo = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(repo / "tools" / "comm_check.py"), "2", "200000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"bit_identical_to_single_gpu_fit_on_every_rank": true' in r.stdout
