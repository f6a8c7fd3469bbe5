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
