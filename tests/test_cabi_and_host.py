"""CPU: the C-ABI library loads and exports every symbol include/avld.h declares; host-only entry points
(pairwise plan, mel taps) agree with numpy / the oracle; the quantile helper reproduces np.quantile;
there is no CPU fallback."""
import ctypes as C
import re

import numpy as np
import pytest

from conftest import REPO
from amphibian_vae_latent_detector_b200 import _lib, quantile
from oracle import librosa_port as lp


def test_library_exports_every_declared_symbol():
    header = (REPO / "include" / "avld.h").read_text()
    declared = set(re.findall(r"\b(avld_[a-z0-9_]+)\s*\(", header))
    declared -= {"avld_ctx", "avld_params", "avld_layer", "avld_rank_query"}
    assert len(declared) >= 20
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in avld.h but not exported"
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    assert lib.avld_abi_version() == 1
    names = [lib.avld_stage_name(i).decode() for i in range(lib.avld_stage_count())]
    assert "prep_kernel" in names and "dftf3_kernel" in names and "fold3_kernel" in names


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = _lib.load()
    h = C.c_void_p()
    p = _lib.Params(48000, 144000, 2048, 384, 64, 150.0, 15000.0, 192, 1e-10, 80.0, 8)
    rc = lib.avld_ctx_create(0, C.byref(p), C.byref(h))
    assert rc == -2 and b"no CPU fallback" in lib.avld_last_error()
    from amphibian_vae_latent_detector_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(0)


def _emulate_pairwise(a, off, ln):
    """float32 evaluation of numpy's pairwise tree from the leaf plan (leaves) + the recursive split."""
    def leaf(v):
        n = len(v)
        if n < 8:
            r = np.float32(0)
            for t in v:
                r = np.float32(r + t)
            return r
        r = v[:8].copy()
        i = 8
        while i < n - n % 8:
            r = (r + v[i:i + 8]).astype(np.float32)
            i += 8
        res = np.float32(np.float32(np.float32(r[0] + r[1]) + np.float32(r[2] + r[3])) +
                         np.float32(np.float32(r[4] + r[5]) + np.float32(r[6] + r[7])))
        for t in v[i:]:
            res = np.float32(res + t)
        return res

    def rec(o, n):
        if n <= 128:
            return leaf(a[o:o + n])
        n2 = n // 2
        n2 -= n2 % 8
        return np.float32(rec(o, n2) + rec(o + n2, n - n2))
    return rec(0, len(a))


@pytest.mark.parametrize("n", [1, 7, 8, 9, 100, 128, 129, 1000, 4097, 144000, 240000])
def test_pairwise_plan_matches_numpy_sum(n):
    lib = _lib.load()
    cap = 1 << 14
    off = (C.c_int64 * cap)()
    ln = (C.c_int64 * cap)()
    nl = lib.avld_pairwise_plan(n, off, ln, cap)
    assert 1 <= nl <= cap
    offs, lens = np.array(off[:nl]), np.array(ln[:nl])
    assert offs[0] == 0 and np.all(offs[1:] == offs[:-1] + lens[:-1]) and offs[-1] + lens[-1] == n
    assert lens.max() <= 128
    if n in (144000, 240000):
        assert nl == 2048 and np.all(lens % 8 == 0)
    a = (np.random.default_rng(n).standard_normal(n).astype(np.float32)) ** 2
    assert _emulate_pairwise(a, offs, lens) == np.add.reduce(a)


def test_mel_taps_match_librosa_port():
    lib = _lib.load()
    p = _lib.Params(48000, 144000, 2048, 384, 64, 150.0, 15000.0, 192, 1e-10, 80.0, 8)
    nb = 1025
    first = (C.c_int32 * nb)()
    w0 = (C.c_float * nb)()
    w1 = (C.c_float * nb)()
    assert lib.avld_mel_taps(C.byref(p), first, w0, w1) == 0
    fb = lp.mel_filterbank(sr=48000, n_fft=2048, n_mels=64, fmin=150.0, fmax=15000.0)      # [64, 1025] float32
    dense = np.zeros_like(fb)
    for b in range(nb):
        if first[b] >= 0:
            dense[first[b], b] = w0[b]
            if w1[b] != 0:
                dense[first[b] + 1, b] = w1[b]
    assert np.count_nonzero(fb) == np.count_nonzero(dense) == 1231
    assert np.max(np.abs(dense - fb)) <= 1e-9 + 2e-7 * np.max(np.abs(fb))
    nz = np.nonzero(fb.sum(axis=0))[0]
    assert nz[0] == 7 and nz[-1] == 640


@pytest.mark.parametrize("n", [1, 2, 3, 7, 100, 399, 1200, 20001])
def test_quantile_helper_matches_numpy(n):
    rng = np.random.default_rng(n)
    x = np.abs(rng.standard_normal(n)).astype(np.float32) * 7
    xs = np.sort(x)
    for q in (0.0, 0.01, 0.10, 0.15, 0.20, 0.25, 0.5, 0.9, 0.95, 1.0):
        prev, nxt, gamma = quantile.neighbour_ranks(n, q, "numpy2")
        got = quantile.lerp(float(xs[prev]), float(xs[min(nxt, n - 1)]), gamma, "numpy2")
        assert got == float(np.quantile(x, q)), (n, q)
        # float64-index variant (numpy 1.26.4, the reference's pin, cannot be executed here): same order
        # statistics, blend within float32 rounding of the float64 quantile
        prev, nxt, gamma = quantile.neighbour_ranks(n, q, "numpy1")
        got1 = quantile.lerp(float(xs[prev]), float(xs[min(nxt, n - 1)]), gamma, "numpy1")
        assert got1 == pytest.approx(float(np.quantile(x.astype(np.float64), q)), rel=1e-6)
