"""GPU: edge cases of the hot path through the C ABI -- empty and ragged batches, batch-composition independence,
NaN / Inf / all-zero / clipped chunks (per-chunk failures never fail the call: 08:504-506, 10:409-418), unsupported
geometries (loud errors, no fallback), other feature parameters against the oracle."""
import numpy as np
import pytest
import torch

from oracle import hotpath as hp
from amphibian_vae_latent_detector_b200 import _lib
from amphibian_vae_latent_detector_b200.engine import Engine, priority_ranks

pytestmark = pytest.mark.gpu
L = 144000
SPECIES = ["Batrachyla_leptopus", "Batrachyla_taeniata", "Calyptocephalella_gayi", "Pleurodema_thaul"]


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-12))


def test_empty_batches(engine3s):
    e = engine3s
    x0 = torch.empty(0, L, device="cuda")
    y, ok, rms = e.rms_normalize(x0)
    assert y.shape == (0, L) and ok.shape == (0,) and rms.shape == (0,)
    assert e.logmel(x0).shape == (0, 192, 64)
    feat, ok, _ = e.normalize_logmel(x0, pcm16=True)
    assert feat.shape == (0, 192, 64)
    mu, ok = e.encode(x0, pcm16=True)
    assert mu.shape == (0, 128)
    assert e.encoder_forward(torch.empty(0, 192, 64, device="cuda")).shape == (0, 128)
    cent = torch.zeros(4, 128, device="cuda")
    r = e.radii(torch.empty(0, 128, device="cuda"), cent)
    assert r.shape == (0, 4)
    pred, best = e.decide(r, torch.ones(4, dtype=torch.float64, device="cuda"), torch.arange(4, dtype=torch.int32, device="cuda"))
    assert pred.shape == (0,) and best.shape == (0,)
    p, b, o, m = e.encode_detect_host(torch.empty(0, L, dtype=torch.int16), np.zeros((4, 128), np.float32), np.ones(4),
                                      np.arange(4, dtype=np.int32), want_mu=True)
    assert p.shape == (0,) and b.shape == (0,) and o.shape == (0,) and m.shape == (0, 128)


def test_ragged_batches_are_batch_independent(engine3s):
    """n = 1, n = max_batch + 1, and the same chunk at different positions of different batches give bit-identical rows."""
    from amphibian_vae_latent_detector_b200 import synth
    x, _ = synth.make_chunks(engine3s.max_batch + 1, L, seed=3, special_every=7)
    X = x.cuda()
    mu_all, ok_all = engine3s.encode(X, pcm16=True)                   # two passes: max_batch + 1
    mu_one, ok_one = engine3s.encode(X[5:6], pcm16=True)
    assert torch.equal(mu_all[5], mu_one[0]) and ok_all[5] == ok_one[0]
    mu_tail, _ = engine3s.encode(X[-3:], pcm16=True)
    assert torch.equal(mu_all[-3:], mu_tail)
    perm = torch.randperm(X.shape[0], generator=torch.Generator().manual_seed(0)).cuda()
    mu_perm, _ = engine3s.encode(X[perm], pcm16=True)
    assert torch.equal(mu_perm, mu_all[perm])


def test_bad_chunks_do_not_poison_their_neighbours(engine3s):
    from amphibian_vae_latent_detector_b200 import synth
    x, _ = synth.make_chunks(8, L, seed=4, special_every=0)
    clean = x.clone()
    x[1, 1000] = float("nan")
    x[2, 2000] = float("inf")
    x[3] = 0.0                                         # digital silence: gate (00:32-34), features all zero (0 / 1e-8)
    x[4] = x[4] * 1e4                                  # absurdly hot: everything clips at +-1 (00:37)
    mu, ok = engine3s.encode(x.cuda(), pcm16=True)
    ref, ok_ref = engine3s.encode(clean.cuda(), pcm16=True)
    mu, ok = mu.cpu().numpy(), ok.cpu().numpy()
    for i in (0, 5, 6, 7):
        assert np.array_equal(mu[i], ref[i].cpu().numpy())
    assert ok[3] == 0 and np.all(np.isfinite(mu[3]))
    yo, oko, _ = hp.rms_normalize_batch(x[3:5].numpy(), pcm16=True)
    assert list(oko) == [ok[3], ok[4]] == [0, 1]
    assert np.all(np.isfinite(mu[4]))
    # the reference's np.mean(y ** 2) of a chunk holding NaN / Inf is NaN / Inf: rms < rms_min is False -> "scaled" with a
    # NaN / 0 scale; what matters here is that the call succeeds and the row is visibly unusable, not silently plausible
    feat, okf, rms = engine3s.normalize_logmel(x[1:3].cuda(), pcm16=True)
    assert not np.isfinite(rms.cpu().numpy()).all()
    # decision on such a row is NO_DETECT (NaN distances accept nothing)
    cent = np.zeros((4, 128), np.float32)
    pred, best, okh, _ = engine3s.encode_detect_host(x[:4].contiguous(), cent, np.full(4, 1e9), priority_ranks(SPECIES, SPECIES))
    assert pred[0] >= 0 and pred[3] >= 0
    assert pred[1] == -1 and pred[2] == -1 and np.isinf(best[1]) and np.isinf(best[2])      # 10:176, :186: min(inf, nan) = inf
    assert np.all(np.isnan(mu[1])) and np.all(np.isnan(mu[2]))


def test_unsupported_geometry_fails_loudly():
    # chunk shorter than the reflect padding: normalisation still works, features refuse (librosa would raise too)
    eng = Engine(0, chunk_len=800, max_batch=4)
    x = torch.randn(2, 800, device="cuda") * 0.01
    y, ok, _ = eng.rms_normalize(x)
    yo, oko, _ = hp.rms_normalize_batch(x.cpu().numpy())
    assert np.array_equal(y.cpu().numpy().view(np.uint32), yo.view(np.uint32))
    with pytest.raises(_lib.AvldError):
        eng.logmel(x)
    eng.close()
    with pytest.raises(ValueError):
        Engine(0, chunk_len=L, max_batch=4).logmel(torch.zeros(2, L))          # CPU tensor: no silent host path
    # more FFT bins with mel weight than the STFT kernel has work items for (12 x 160 bins per class split): refused at
    # context creation, and so is an n_fft the folded GEMM cannot tile
    with pytest.raises(_lib.AvldError) as e:
        Engine(0, chunk_len=L, max_batch=4, sr=48000, n_fft=8192, fmin=0.0, fmax=24000.0)
    assert e.value.code == -3                       # AVLD_ERR_UNSUPPORTED
    with pytest.raises(_lib.AvldError) as e:
        Engine(0, chunk_len=L, max_batch=4, n_fft=1000, hop_length=250)
    assert e.value.code == -3


@pytest.mark.parametrize("kw", [
    dict(sr=48000, n_fft=1024, hop_length=256, n_mels=40, fmin=300.0, fmax=12000.0, target_frames=128),
    dict(sr=32000, n_fft=2048, hop_length=512, n_mels=64, fmin=50.0, fmax=11000.0, target_frames=96),
    dict(sr=48000, n_fft=2048, hop_length=384, n_mels=128, fmin=150.0, fmax=15000.0, target_frames=192),
    dict(sr=32000, n_fft=2048, hop_length=512, n_mels=62, fmin=50.0, fmax=14000.0, target_frames=96),   # 7 work items; n_mels % 4 != 0: scalar reductions
])
def test_other_feature_parameters_vs_oracle(kw):
    """The kernels are not specialised to the CLI defaults: other n_fft / hop / n_mels / band limits against the oracle."""
    from amphibian_vae_latent_detector_b200 import synth
    Lc = 96000
    x, _ = synth.make_chunks(5, Lc, sr=kw["sr"], seed=21, special_every=0)
    eng = Engine(0, chunk_len=Lc, max_batch=4, **kw)
    feat, ok, _ = eng.normalize_logmel(x.cuda(), pcm16=True)
    yo, oko, _ = hp.rms_normalize_batch(x.numpy(), pcm16=True)
    fo = hp.logmel_features_batch(yo, **kw)
    assert np.array_equal(ok.cpu().numpy(), oko)
    errs = [rel(feat[i].cpu().numpy(), fo[i]) for i in range(len(fo))]
    assert max(errs) < 2e-4, (eng.dft_info()["mode"], errs)
    eng.close()


def test_outputs_do_not_overrun_their_buffers(engine3s):
    """Every caller-owned output lives inside a larger allocation with sentinel bytes on both sides (the pool has no
    compute-sanitizer: this is the bounds check for writes).  Odd n so that tile tails are exercised."""
    from amphibian_vae_latent_detector_b200 import synth
    e, n, pad = engine3s, 37, 4096
    x, lab = synth.make_chunks(n, L, seed=12, special_every=9)
    X = x.cuda()

    def guarded(shape, dtype):
        numel = int(np.prod(shape))
        big = torch.full((numel + 2 * pad,), 77, dtype=dtype, device="cuda")
        return big, big[pad:pad + numel].view(*shape)

    def intact(big, numel):
        return bool((big[:pad] == 77).all() and (big[pad + numel:] == 77).all())

    lib, h, st = e.lib, e._h, torch.cuda.current_stream().cuda_stream
    big_y, y = guarded((n, L), torch.float32)
    big_ok, ok = guarded((n,), torch.uint8)
    big_rms, rms = guarded((n,), torch.float32)
    _lib.check(lib.avld_rms_normalize(h, X.data_ptr(), y.data_ptr(), ok.data_ptr(), rms.data_ptr(), n, 0.05, 1e-4, 1e-8, 1, st))
    big_f, feat = guarded((n, 192, 64), torch.float32)
    _lib.check(lib.avld_normalize_logmel(h, X.data_ptr(), feat.data_ptr(), ok.data_ptr(), rms.data_ptr(), n, 0.05, 1e-4, 1e-8, 1, st))
    big_mu, mu = guarded((n, 128), torch.float32)
    _lib.check(lib.avld_encoder_forward(h, feat.data_ptr(), mu.data_ptr(), n, st))
    big_mu2, mu2 = guarded((n, 128), torch.float32)
    _lib.check(lib.avld_encode(h, X.data_ptr(), mu2.data_ptr(), ok.data_ptr(), n, 0.05, 1e-4, 1e-8, 1, st))
    cent = torch.randn(4, 128, device="cuda")
    big_r, radii = guarded((n, 4), torch.float32)
    _lib.check(lib.avld_radii(h, mu.data_ptr(), cent.data_ptr(), radii.data_ptr(), n, 4, 128, st))
    big_p, pred = guarded((n,), torch.int32)
    big_b, best = guarded((n,), torch.float32)
    thr = torch.full((4,), 50.0, dtype=torch.float64, device="cuda")
    prio = torch.arange(4, dtype=torch.int32, device="cuda")
    _lib.check(lib.avld_decide(h, radii.data_ptr(), thr.data_ptr(), prio.data_ptr(), pred.data_ptr(), best.data_ptr(), n, 4, st))
    torch.cuda.synchronize()
    for name, big, numel in (("y", big_y, n * L), ("ok", big_ok, n), ("rms", big_rms, n), ("feat", big_f, n * 192 * 64),
                             ("mu", big_mu, n * 128), ("mu2", big_mu2, n * 128), ("radii", big_r, n * 4),
                             ("pred", big_p, n), ("best", big_b, n)):
        assert intact(big, numel), name
    assert torch.equal(mu, mu2)
    assert not bool((y == 77).all()) and not bool((feat == 77).any())
