"""GPU: avld_rms_normalize is bit-exact against the reference outputs in tests/golden (made by the
reference's own rms_normalize, 00_normalize_dataset_rms.py:29-38) and against the numpy oracle."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_pcm_case
from oracle import hotpath as hp
from oracle import librosa_port as lp

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("length,fixture", [(144000, "engine3s"), (240000, "engine5s")])
def test_rms_bit_exact_golden(length, fixture, golden_meta, request):
    eng = request.getfixturevalue(fixture)
    keys = [k for k in golden_meta["rms_cases"] if k.endswith(f"_{length}")]
    assert len(keys) >= 7
    X = np.stack([load_pcm_case(GOLDEN / f"rms_{k}.npz")[0] for k in keys])
    y, ok, rms = eng.rms_normalize(torch.from_numpy(X).cuda())
    y, ok, rms = y.cpu().numpy(), ok.cpu().numpy(), rms.cpu().numpy()
    for i, k in enumerate(keys):
        info = golden_meta["rms_cases"][k]
        assert bool(ok[i]) == info["ok"], k
        assert sha(y[i]) == info["sha"], k                      # every sample, bit for bit
        assert rms[i] == np.float32(info["rms"]), k


def test_rms_bit_exact_random_batch(engine3s):
    from amphibian_vae_latent_detector_b200 import synth
    x, _ = synth.make_chunks(48, 144000, seed=5, special_every=10)
    yo, oko, rmso = hp.rms_normalize_batch(x.numpy())
    y, ok, rms = engine3s.rms_normalize(x.cuda())
    assert np.array_equal(ok.cpu().numpy(), oko) and 0 < oko.sum() < len(oko)
    assert np.array_equal(rms.cpu().numpy(), rmso)
    assert np.array_equal(y.cpu().numpy().view(np.uint32), yo.view(np.uint32))
    # the sf.write(PCM_16) + librosa.load round trip of process_folder (00:55-57)
    yq, _, _ = engine3s.rms_normalize(x.cuda(), pcm16=True)
    yqo, _, _ = hp.rms_normalize_batch(x.numpy(), pcm16=True)
    assert np.array_equal(yq.cpu().numpy().view(np.uint32), yqo.view(np.uint32))


def test_rms_odd_lengths():
    """chunk lengths that are not multiples of 8 / below one leaf: the tail and n<8 branches of numpy's pairwise sum."""
    from amphibian_vae_latent_detector_b200.engine import Engine
    rng = np.random.default_rng(3)
    for L in (1031, 4097, 5000, 130001):   # > n_fft/2 (reflect padding of the feature stage)
        eng = Engine(0, chunk_len=L, max_batch=4, n_fft=2048, hop_length=1024, target_frames=8)
        x = (0.1 * rng.standard_normal((5, L))).astype(np.float32)
        x[3] *= 1e-4
        yo, oko, rmso = hp.rms_normalize_batch(x)
        y, ok, rms = eng.rms_normalize(torch.from_numpy(x).cuda())
        assert np.array_equal(ok.cpu().numpy(), oko), L
        assert np.array_equal(rms.cpu().numpy(), rmso), L
        assert np.array_equal(y.cpu().numpy().view(np.uint32), yo.view(np.uint32)), L
        eng.close()


def test_rms_properties_full_size(engine3s):
    """Size-independent properties on a large batch: rms(y) == 0.05 when nothing clipped; gate leaves x untouched; idempotence of the gate."""
    from amphibian_vae_latent_detector_b200 import synth
    x, _ = synth.make_chunks(512, 144000, seed=9, device="cuda")
    y, ok, rms = engine3s.rms_normalize(x)
    okb = ok.bool()
    assert torch.equal(y[~okb], x[~okb])
    unclipped = okb & (y.abs().amax(dim=1) < 1.0)
    r = y[unclipped].double().pow(2).mean(dim=1).sqrt()
    assert unclipped.sum() > 400 and float((r - 0.05).abs().max()) < 1e-6
    assert float(y.abs().max()) <= 1.0


def test_rms_numpy1_scalar_semantics_bit_exact(standin_encoder):
    """scalar_semantics="numpy1": `rms + eps`, `0.05 / (...)` and the gate in float64 as the reference's pinned numpy==1.26.4
    evaluates them (00_normalize_dataset_rms.py:30-36) -- bit-exact against the oracle's numpy-1 variant, different from the
    numpy-2 result in a sizeable share of the chunks (one ulp of the scale), and carried through the fused paths."""
    from amphibian_vae_latent_detector_b200 import synth
    from amphibian_vae_latent_detector_b200.engine import Engine
    x, _ = synth.make_chunks(96, 144000, seed=41, special_every=10)
    xn = x.numpy()
    eng1 = Engine(0, chunk_len=144000, max_batch=32, scalar_semantics="numpy1")
    y1, ok1, rms1 = eng1.rms_normalize(x.cuda())
    y1o, ok1o, rms1o = hp.rms_normalize_batch(xn, numpy1_scalars=True)
    assert np.array_equal(ok1.cpu().numpy(), ok1o) and np.array_equal(rms1.cpu().numpy(), rms1o)
    assert np.array_equal(y1.cpu().numpy().view(np.uint32), y1o.view(np.uint32))
    y2o, _, _ = hp.rms_normalize_batch(xn)
    differ = np.any(y1o.view(np.uint32) != y2o.view(np.uint32), axis=1)
    assert 0.1 < differ.mean() < 0.9, differ.mean()            # the two numpy generations really disagree on many chunks
    # PCM_16 round trip and the feature / encode paths follow the same scale
    y1q, _, _ = eng1.rms_normalize(x.cuda(), pcm16=True)
    y1qo, _, _ = hp.rms_normalize_batch(xn, pcm16=True, numpy1_scalars=True)
    assert np.array_equal(y1q.cpu().numpy().view(np.uint32), y1qo.view(np.uint32))
    f_direct = eng1.logmel(y1q)                                 # features of the already normalised audio
    f_fused, okf, _ = eng1.normalize_logmel(x.cuda(), pcm16=True)
    assert np.array_equal(okf.cpu().numpy(), ok1o)
    assert torch.equal(f_direct, f_fused)
    # non-default constants travel through the context too (fused host path included)
    eng1.close()
    eng3 = Engine(0, chunk_len=144000, max_batch=32, scalar_semantics="numpy1", target_rms=0.1, rms_min=2e-3, eps=1e-6)
    y3, ok3, _ = eng3.rms_normalize(x.cuda(), target_rms=0.1, rms_min=2e-3, eps=1e-6)
    y3o, ok3o, _ = hp.rms_normalize_batch(xn, 0.1, 2e-3, 1e-6, numpy1_scalars=True)
    assert np.array_equal(ok3.cpu().numpy(), ok3o)
    assert np.array_equal(y3.cpu().numpy().view(np.uint32), y3o.view(np.uint32))
    from amphibian_vae_latent_detector_b200 import _lib
    with pytest.raises(_lib.AvldError):                          # per-call constants that contradict the context's
        eng3.rms_normalize(x.cuda(), target_rms=0.05)
    y3b, _, _ = eng3.rms_normalize(x.cuda())                      # defaults = the engine's own constants
    assert torch.equal(y3b, y3)
    eng3.load_encoder(standin_encoder)
    cent = np.zeros((4, 128), np.float32)
    thr = np.full(4, 1e9)
    from amphibian_vae_latent_detector_b200.engine import priority_ranks
    _, _, ok_h, _ = eng3.encode_detect_host(x[:40].pin_memory(), cent, thr, priority_ranks(hp.PRIORITY_ORDER, hp.PRIORITY_ORDER), pcm16=True)
    assert np.array_equal(ok_h, ok3o[:40])
    eng3.close()
