"""GPU: whole path through the host-buffer C-ABI call (avld_encode_detect_host): raw chunks on the host ->
decisions, vs the numpy oracle run chunk by chunk the way the reference does (config 1, reduced to 64 chunks
so the CPU side finishes in seconds)."""
import numpy as np
import pytest
import torch

from oracle import hotpath as hp
from amphibian_vae_latent_detector_b200 import synth
from amphibian_vae_latent_detector_b200.engine import priority_ranks

pytestmark = pytest.mark.gpu
MEL_KW = dict(sr=48000, n_mels=64, fmin=150.0, fmax=15000.0, hop_length=384, n_fft=2048, target_frames=192)
SPECIES = ["Batrachyla_leptopus", "Batrachyla_taeniata", "Calyptocephalella_gayi", "Pleurodema_thaul"]


def test_encode_fit_detect_vs_oracle(engine3s, standin_encoder):
    n = 96
    x, label = synth.make_chunks(n, 144000, seed=123, special_every=25)
    # ---- oracle, per chunk (the reference's execution model)
    yo, oko, _ = hp.rms_normalize_batch(x.numpy(), pcm16=True)
    Zo = hp.encode_batch(standin_encoder, yo, **MEL_KW)
    cent_o, rk_o, rk_in_o, rk_out_o = hp.fit_radial(Zo, label.numpy(), 4, 0.95, 0.10)
    pred_o, best_o, radii_o = hp.decide_batch(Zo, SPECIES, cent_o, rk_o)

    # ---- GPU: encode on device, fit, then the host-buffer end-to-end detect call
    Z, ok = engine3s.encode(x.cuda(), pcm16=True)
    assert np.array_equal(ok.cpu().numpy(), oko)
    Zn = Z.cpu().numpy()
    assert np.max(np.abs(Zn - Zo)) / np.max(np.abs(Zo)) < 1e-3
    fit = engine3s.fit_radial(Z, label.cuda(), 4, 0.95, 0.10)
    assert np.max(np.abs(fit.centroids - cent_o)) / np.max(np.abs(cent_o)) < 1e-3
    assert np.allclose(fit.rk[0], rk_o, rtol=1e-3)
    prio = priority_ranks(SPECIES, hp.PRIORITY_ORDER)
    xp = x.pin_memory()
    pred, best, ok_h, mu_h = engine3s.encode_detect_host(xp, cent_o, rk_o, prio, pcm16=True, want_mu=True)
    assert np.array_equal(ok_h, oko)
    assert np.max(np.abs(mu_h - Zo)) / np.max(np.abs(Zo)) < 1e-3
    # decisions identical except chunks within 1e-3 of a threshold
    near = np.any(np.abs(radii_o - rk_o[None]) / rk_o[None] <= 1e-3, axis=1)
    assert np.array_equal(pred[~near], pred_o[~near]), (pred, pred_o)
    assert near.sum() < n // 4
    assert np.allclose(best, best_o, rtol=1e-3)
    assert len(set(pred_o.tolist())) >= 3          # detections of several species and NO_DETECT all occur


def test_pcm16_host_input_equals_float_input(engine3s):
    """int16 PCM in (decoded on the GPU as s/32768, librosa.load semantics) == the same samples passed as float32."""
    x, label = synth.make_chunks(70, 144000, seed=5, special_every=9)
    pcm = torch.clamp(torch.round(x * 20000.0), -32768, 32767).to(torch.int16)
    xf = pcm.to(torch.float32) * (1.0 / 32768.0)
    rng = np.random.default_rng(0)
    cent = rng.standard_normal((4, 128)).astype(np.float32)
    thr = np.array([30.0, 40.0, 50.0, 60.0])
    prio = priority_ranks(SPECIES, hp.PRIORITY_ORDER)
    a = engine3s.encode_detect_host(pcm.pin_memory(), cent, thr, prio, pcm16=True, want_mu=True)
    b = engine3s.encode_detect_host(xf.pin_memory(), cent, thr, prio, pcm16=True, want_mu=True)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)


def test_device_pcm16_encode_equals_float_encode(engine3s):
    """avld_encode_pcm16 (device-resident PCM_16) == avld_encode on the decoded float32 samples, bit for bit."""
    x, _ = synth.make_chunks(70, 144000, seed=6, special_every=9)
    pcm = torch.clamp(torch.round(x * 20000.0), -32768, 32767).to(torch.int16)
    xf = pcm.to(torch.float32) * (1.0 / 32768.0)
    mu_a, ok_a = engine3s.encode(pcm.cuda(), pcm16=True)
    mu_b, ok_b = engine3s.encode(xf.cuda(), pcm16=True)
    assert torch.equal(mu_a, mu_b) and torch.equal(ok_a, ok_b)


def test_config_c1_1000_chunks_vs_oracle(standin_encoder):
    """BASELINE.json configs[0] at full size: 1 000 synthetic chunks, normalise + encode + radial fit over the q_out grid
    + decision; CPU oracle (one chunk at a time, as the reference runs) against the batched CUDA path."""
    from concurrent.futures import ThreadPoolExecutor
    from amphibian_vae_latent_detector_b200.engine import Engine
    n, grid = 1000, (0.10, 0.15, 0.20, 0.25)
    x, label = synth.make_chunks(n, 144000, seed=123, special_every=100)
    xn, ln = x.numpy(), label.numpy()
    yo, oko, _ = hp.rms_normalize_batch(xn, pcm16=True)
    threads = torch.get_num_threads()
    torch.set_num_threads(1)                       # the oracle pieces run one chunk per thread (numpy releases the GIL)
    try:
        with ThreadPoolExecutor(16) as ex:
            parts = list(ex.map(lambda i: hp.encode_batch(standin_encoder, yo[i:i + 25], **MEL_KW), range(0, n, 25)))
    finally:
        torch.set_num_threads(threads)
    Zo = np.concatenate(parts)
    eng = Engine(0, chunk_len=144000, max_batch=256)
    eng.load_encoder(standin_encoder)
    Z, ok = eng.encode(x.cuda(), pcm16=True)
    assert np.array_equal(ok.cpu().numpy(), oko)
    assert np.max(np.abs(Z.cpu().numpy() - Zo)) / np.max(np.abs(Zo)) < 1e-3
    fit = eng.fit_radial(Z, label.cuda(), 4, 0.95, grid)
    prio = priority_ranks(SPECIES, hp.PRIORITY_ORDER)
    flips = 0
    for qi, q in enumerate(grid):
        cent_o, rk_o, rk_in_o, rk_out_o = hp.fit_radial(Zo, ln, 4, 0.95, q)
        assert np.max(np.abs(fit.centroids - cent_o)) / np.max(np.abs(cent_o)) < 1e-3
        assert np.allclose(fit.rk_in, rk_in_o, rtol=1e-3) and np.allclose(fit.rk_out[qi], rk_out_o, rtol=1e-3)
        assert np.allclose(fit.rk[qi], rk_o, rtol=1e-3)
        pred_o, best_o, radii_o = hp.decide_batch(Zo, SPECIES, cent_o, rk_o)
        pred, best = eng.decide(fit.radii_local, torch.from_numpy(fit.rk[qi]).cuda(), torch.from_numpy(prio).cuda())
        pred, best = pred.cpu().numpy(), best.cpu().numpy()
        near = np.any(np.abs(radii_o - rk_o[None]) / rk_o[None] <= 1e-3, axis=1)      # north star: identical outside 1e-3 of a threshold
        assert np.array_equal(pred[~near], pred_o[~near])
        assert np.allclose(best, best_o, rtol=1e-3)
        flips += int((pred != pred_o).sum())
        assert near.sum() < n // 10
    assert flips <= n // 50
    eng.close()


def test_5s_chunks_latents_and_decisions_vs_oracle(standin_encoder):
    """chunk_seconds = 5.0, the default of 08 / 09 / 10 (08:392-396, core:358-370): L = 240 000 samples, 626 frames, centre crop
    at frame 217 -- latents, radial fit and decisions against the oracle, float32 and PCM_16 input, host-buffer path included."""
    from amphibian_vae_latent_detector_b200.engine import Engine
    n, L5 = 96, 240000
    x, label = synth.make_chunks(n, L5, seed=77, special_every=11)
    xn, ln = x.numpy(), label.numpy()
    yo, oko, _ = hp.rms_normalize_batch(xn, pcm16=True)
    Zo = hp.encode_batch(standin_encoder, yo, **MEL_KW)
    eng = Engine(0, chunk_len=L5, max_batch=32)
    eng.load_encoder(standin_encoder)
    assert eng.n_frames == 626
    Z, ok = eng.encode(x.cuda(), pcm16=True)
    assert np.array_equal(ok.cpu().numpy(), oko)
    assert np.max(np.abs(Z.cpu().numpy() - Zo)) / np.max(np.abs(Zo)) < 1e-3
    fit = eng.fit_radial(Z, label.cuda(), 4, 0.95, (0.25,))
    cent_o, rk_o, _, _ = hp.fit_radial(Zo, ln, 4, 0.95, 0.25)
    assert np.allclose(fit.rk[0], rk_o, rtol=1e-3) and np.max(np.abs(fit.centroids - cent_o)) / np.max(np.abs(cent_o)) < 1e-3
    prio = priority_ranks(SPECIES, hp.PRIORITY_ORDER)
    pred_o, best_o, radii_o = hp.decide_batch(Zo, SPECIES, cent_o, rk_o)
    near = np.any(np.abs(radii_o - rk_o[None]) / rk_o[None] <= 1e-3, axis=1)
    pred, best = eng.decide(fit.radii_local, torch.from_numpy(fit.rk[0]).cuda(), torch.from_numpy(prio).cuda())
    assert np.array_equal(pred.cpu().numpy()[~near], pred_o[~near]) and np.allclose(best.cpu().numpy(), best_o, rtol=1e-3)
    # the host-buffer call on PCM_16 samples, as the 5 s WAV files of the reference's datasets hold them
    pcm = torch.clamp(torch.round(x * 32767.0), -32768, 32767).to(torch.int16)
    yq, okq, _ = hp.rms_normalize_batch(pcm.numpy().astype(np.float32) / 32768.0, pcm16=True)
    Zq = hp.encode_batch(standin_encoder, yq, **MEL_KW)
    pred_h, best_h, ok_h, mu_h = eng.encode_detect_host(pcm.pin_memory(), cent_o, rk_o, prio, pcm16=True, want_mu=True)
    assert np.array_equal(ok_h, okq) and np.max(np.abs(mu_h - Zq)) / np.max(np.abs(Zq)) < 1e-3
    pq, bq, rq = hp.decide_batch(Zq, SPECIES, cent_o, rk_o)
    nearq = np.any(np.abs(rq - rk_o[None]) / rk_o[None] <= 1e-3, axis=1)
    assert np.array_equal(pred_h[~nearq], pq[~nearq])
    eng.close()
