"""CPU: numpy emulation of the folded STFT operands (tools/split_precision_study.py).  Two facts are pinned:
the three-level fold + fp16 hi/lo splits in three passes reproduce the reference-made features (so the emulation is the
algorithm the kernels implement), and dropping either correction pass breaks the 1e-3 latent tolerance on a tonal chunk
(so the third pass is not optional)."""
import importlib.util
from pathlib import Path

import numpy as np

from conftest import GOLDEN, load_pcm_case
from oracle import hotpath as hp
from oracle import librosa_port as lp

TOOL = Path(__file__).resolve().parents[1] / "tools" / "split_precision_study.py"


def test_three_passes_are_needed_and_sufficient(standin_encoder):
    spec = importlib.util.spec_from_file_location("split_precision_study", TOOL)
    st = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(st)
    fb = lp.mel_filterbank(sr=48000, n_fft=2048, n_mels=64, fmin=150.0, fmax=15000.0)
    bins = np.nonzero(fb.sum(axis=0) > 0)[0]
    assert bins.shape[0] == 634                                        # the bins with mel weight
    x, d = load_pcm_case(GOLDEN / "feat_tonal_3s.npz")
    y, _, _ = hp.rms_normalize_batch(x[None], pcm16=True)
    u = st.frames_of(y[0])
    err = {}
    for v in ("exact", "3pass", "noBlo"):
        feat = st.features_from_power(st.spectrum(u, v, bins), bins, fb)
        mu = hp.encode_features(standin_encoder, feat)
        err[v] = (float(np.max(np.abs(feat - d["feat"])) / np.max(np.abs(d["feat"]))),
                  float(np.max(np.abs(mu - d["z"])) / np.max(np.abs(d["z"]))))
    assert err["exact"][0] < 5e-5 and err["3pass"][0] < 5e-5 and err["3pass"][1] < 2e-5
    assert err["noBlo"][0] > 2e-3 and err["noBlo"][1] > 1e-3
