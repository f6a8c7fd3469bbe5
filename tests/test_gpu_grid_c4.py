"""GPU: BASELINE.json configs[3] -- the q_out calibration grid (9200:20, run_qout_grid.sh:6-13) over 10 M precomputed
latents with multi-species centroids, at full size.  The oracle cannot run at this size in seconds, so the full-size
checks are properties plus an independent GPU computation (torch.sort) of the same order statistics; the small-size
cases of the same code path are pinned to the reference in test_gpu_radial.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GRID = (0.10, 0.15, 0.20, 0.25)


def _device_latents(n, d, k, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    cents = 3.0 * torch.randn(k, d, generator=g, device="cuda")
    label = (torch.arange(n, device="cuda") % k).to(torch.int32)
    Z = torch.empty(n, d, dtype=torch.float32, device="cuda")
    step = 1 << 20
    for i in range(0, n, step):
        m = min(step, n - i)
        Z[i:i + m] = cents[label[i:i + m].long()] + torch.randn(m, d, generator=g, device="cuda")
    return Z, label, cents


def _np_quantile_from_sorted(s: torch.Tensor, q: float) -> float:
    """np.quantile(x, q) (linear, H&F 7) from a sorted float32 vector, in float64 like numpy's scalar path."""
    n = s.shape[0]
    pos = q * (n - 1)
    lo = int(np.floor(pos))
    hi = min(lo + 1, n - 1)
    a, b = float(s[lo].item()), float(s[hi].item())
    t = pos - lo
    return a + (b - a) * t


@pytest.mark.parametrize("n,d,k", [(10_000_000, 128, 4), (2_000_000, 128, 32)])
def test_grid_full_size(engine3s, n, d, k):
    Z, label, cents = _device_latents(n, d, k, 123)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    engine3s.fit_radial(Z[:100000], label[:100000], k, 0.95, GRID)      # warm-up
    ev0.record()
    fit = engine3s.fit_radial(Z, label, k, 0.95, GRID)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    print(f"\nC4 grid: n={n} K={k} D={d}: fit over {len(GRID)} q_out in {ms:.1f} ms "
          f"({n / ms / 1e3:.1f} M latents/s, {4.0 * n * d / ms / 1e6:.0f} GB/s of latents)")
    # centroids: float64 sums (08:316 is np.mean in float32 pairwise; 1e-6 is far inside the 1e-3 budget)
    for kk in (0, k - 1):
        ref = Z[label == kk].double().mean(dim=0).float().cpu().numpy()
        assert np.max(np.abs(fit.centroids[kk] - ref)) <= 1e-5 * max(1.0, np.max(np.abs(ref)))
    assert int(fit.counts.sum()) == n
    # thresholds: monotone in q_out, rk = min(rk_in, rk_out), and equal to an independent sort-based quantile
    assert np.all(np.diff(fit.rk_out, axis=0) >= 0)
    assert np.array_equal(fit.rk, np.minimum(fit.rk_in[None, :], fit.rk_out))
    r = fit.radii_local
    for kk in (0, k - 1):
        s_in = torch.sort(r[label == kk, kk]).values
        s_out = torch.sort(r[label != kk, kk]).values
        assert fit.rk_in[kk] == pytest.approx(_np_quantile_from_sorted(s_in, 0.95), rel=1e-6)
        for qi, q in enumerate(GRID):
            assert fit.rk_out[qi, kk] == pytest.approx(_np_quantile_from_sorted(s_out, q), rel=1e-6)
        assert fit.summaries["in"][kk, 0] == pytest.approx(float(s_in[0]), rel=1e-6)
        assert fit.summaries["out"][kk, 3] == pytest.approx(float(s_out[-1]), rel=1e-6)
    # decisions for all four threshold sets in one pass over the cached radii: counts are monotone in q_out
    from amphibian_vae_latent_detector_b200.engine import priority_ranks
    prio = torch.from_numpy(priority_ranks([f"s{i:02d}" for i in range(k)], [])).cuda()
    detected = []
    for qi in range(len(GRID)):
        pred, best = engine3s.decide(r, torch.from_numpy(fit.rk[qi]).cuda(), prio)
        detected.append(int((pred >= 0).sum()))
        acc_rows = (pred >= 0)
        assert torch.all(best[acc_rows] <= float(np.max(fit.rk[qi])) * (1 + 1e-6))
    assert detected == sorted(detected)
    del Z, r
    torch.cuda.empty_cache()
