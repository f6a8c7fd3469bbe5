"""TEST INFRASTRUCTURE ONLY -- restatement of the third-party arithmetic on the hot path.

The reference calls, but does not vendor (requirements-thesis-baseline-macos-arm64.txt:33,:81):

* ``librosa==0.9.2``: ``librosa.load`` (00_normalize_dataset_rms.py:51, map_detector_core.py:210),
  ``librosa.feature.melspectrogram`` (map_detector_core.py:219-228),
  ``librosa.power_to_db`` (map_detector_core.py:229).
* ``soundfile==0.13.1``: ``sf.write`` (00_normalize_dataset_rms.py:57).

Neither package is installed in this image, so their *published* algorithms are restated
here in numpy, following the librosa 0.9.2 sources (``core/spectrum.py::stft``,
``_spectrogram``, ``power_to_db``; ``filters.py::mel``; ``core/convert.py::hz_to_mel`` ...)
and libsndfile's float<->PCM_16 conversion (``f2s_array`` scales by 0x7FFF with ``lrintf``;
``s2f_array`` divides by 0x8000).  Parity vs the real librosa is unpinned (no copy, no
golden vectors in the reference); the restatement is pinned against torchaudio's
independent implementation in tests/test_oracle_pinning.py.
"""
from __future__ import annotations

import wave
from pathlib import Path
from typing import Callable, Optional, Tuple, Union

import numpy as np
import scipy.signal

# ----------------------------------------------------------------------------------------
# librosa.core.convert
# ----------------------------------------------------------------------------------------
_F_MIN = 0.0
_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = (_MIN_LOG_HZ - _F_MIN) / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(frequencies, htk: bool = False):
    """Slaney (htk=False) / HTK mel scale, librosa 0.9.2 ``hz_to_mel``."""
    frequencies = np.asanyarray(frequencies, dtype=np.float64)
    if htk:
        return 2595.0 * np.log10(1.0 + frequencies / 700.0)
    mels = (frequencies - _F_MIN) / _F_SP
    if frequencies.ndim:
        log_t = frequencies >= _MIN_LOG_HZ
        mels[log_t] = _MIN_LOG_MEL + np.log(frequencies[log_t] / _MIN_LOG_HZ) / _LOGSTEP
    elif frequencies >= _MIN_LOG_HZ:
        mels = _MIN_LOG_MEL + np.log(frequencies / _MIN_LOG_HZ) / _LOGSTEP
    return mels


def mel_to_hz(mels, htk: bool = False):
    mels = np.asanyarray(mels, dtype=np.float64)
    if htk:
        return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    freqs = _F_MIN + _F_SP * mels
    if mels.ndim:
        log_t = mels >= _MIN_LOG_MEL
        freqs[log_t] = _MIN_LOG_HZ * np.exp(_LOGSTEP * (mels[log_t] - _MIN_LOG_MEL))
    elif mels >= _MIN_LOG_MEL:
        freqs = _MIN_LOG_HZ * np.exp(_LOGSTEP * (mels - _MIN_LOG_MEL))
    return freqs


def fft_frequencies(sr: float = 22050, n_fft: int = 2048) -> np.ndarray:
    return np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)


def mel_frequencies(n_mels: int = 128, fmin: float = 0.0, fmax: float = 11025.0, htk: bool = False):
    min_mel = hz_to_mel(fmin, htk=htk)
    max_mel = hz_to_mel(fmax, htk=htk)
    mels = np.linspace(min_mel, max_mel, n_mels)
    return mel_to_hz(mels, htk=htk)


# ----------------------------------------------------------------------------------------
# librosa.filters.mel
# ----------------------------------------------------------------------------------------
def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney",
                   dtype=np.float32) -> np.ndarray:
    """``librosa.filters.mel`` (0.9.2): float64 triangles cast row-wise into a ``dtype`` array,
    then slaney area normalisation applied in place."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=dtype)
    fftfreqs = fft_frequencies(sr=sr, n_fft=n_fft)
    mel_f = mel_frequencies(n_mels + 2, fmin=fmin, fmax=fmax, htk=htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if norm == "slaney":
        enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
        weights *= enorm[:, np.newaxis]
    elif norm is not None:
        raise ValueError(f"unsupported norm={norm!r} in the oracle port")
    return weights


# ----------------------------------------------------------------------------------------
# librosa.core.spectrum
# ----------------------------------------------------------------------------------------
def _pad_center(data: np.ndarray, size: int) -> np.ndarray:
    n = data.shape[-1]
    lpad = int((size - n) // 2)
    return np.pad(data, (lpad, int(size - n - lpad)), mode="constant")


def stft(y: np.ndarray, n_fft=2048, hop_length=None, win_length=None, window="hann",
         center=True, pad_mode="reflect") -> np.ndarray:
    """``librosa.stft`` 0.9.2 for 1-D ``y``: periodic window in float64, reflect padding of
    ``n_fft//2`` each side, frames ``[n_fft, 1 + len(y)//hop]``, ``np.fft.rfft`` of the float64
    product, result stored as complex64 (``util.dtype_r2c(float32)``)."""
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    fft_window = scipy.signal.get_window(window, win_length, fftbins=True)
    fft_window = _pad_center(fft_window, n_fft).reshape((-1, 1))
    y = np.asarray(y)
    if center:
        if n_fft > y.shape[-1]:
            pass  # librosa only warns
        y = np.pad(y, (int(n_fft // 2), int(n_fft // 2)), mode=pad_mode)
    n_frames = 1 + (y.shape[-1] - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(n_frames)[None, :]
    y_frames = y[idx]                                   # [n_fft, n_frames], dtype of y
    out_dtype = np.complex64 if y.dtype == np.float32 else np.complex128
    stft_matrix = np.empty((1 + n_fft // 2, n_frames), dtype=out_dtype, order="F")
    # librosa processes column blocks of <= MAX_MEM_BLOCK bytes; the arithmetic per column is
    # independent, so block boundaries do not change any value.
    block = max(1, 2 ** 18 // (stft_matrix.shape[0] * stft_matrix.itemsize))
    for s in range(0, n_frames, block):
        t = min(s + block, n_frames)
        stft_matrix[:, s:t] = np.fft.rfft(fft_window * y_frames[:, s:t], axis=0)
    return stft_matrix


def melspectrogram(y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None,
                   window="hann", center=True, pad_mode="reflect", power=2.0, **kwargs):
    """``librosa.feature.melspectrogram`` 0.9.2: ``S = |stft|**power`` (float32 for float32 audio),
    ``mel_basis = filters.mel(sr, n_fft, **kwargs)``, ``einsum('...ft,mf->...mt')``."""
    if S is None:
        S = np.abs(stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                        window=window, center=center, pad_mode=pad_mode)) ** power
    mel_basis = mel_filterbank(sr=sr, n_fft=n_fft, **kwargs)
    return np.einsum("...ft,mf->...mt", S, mel_basis, optimize=True)


def power_to_db(S, ref: Union[float, Callable] = 1.0, amin: float = 1e-10,
                top_db: Optional[float] = 80.0) -> np.ndarray:
    """``librosa.power_to_db`` 0.9.2."""
    S = np.asarray(S)
    if amin <= 0:
        raise ValueError("amin must be strictly positive")
    magnitude = np.abs(S) if np.issubdtype(S.dtype, np.complexfloating) else S
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        if top_db < 0:
            raise ValueError("top_db must be non-negative")
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


# ----------------------------------------------------------------------------------------
# WAV I/O: soundfile.write (float -> PCM_16) and librosa.load (PCM -> float32 mono)
# ----------------------------------------------------------------------------------------
def float_to_pcm16(y: np.ndarray) -> np.ndarray:
    """libsndfile ``f2s_array`` with normalisation on and clipping off: ``lrintf(x * 0x7FFF)``
    (float32 product, round-half-even).  Inputs are already inside [-1, 1] on this path."""
    y = np.asarray(y, dtype=np.float32)
    return np.rint(y * np.float32(32767.0)).astype(np.int16)


def pcm16_to_float(s: np.ndarray) -> np.ndarray:
    """libsndfile ``s2f_array``: ``x = s / 0x8000`` in float32 (exact)."""
    return (np.asarray(s, dtype=np.int16).astype(np.float32)) * np.float32(1.0 / 32768.0)


def pcm16_roundtrip(y: np.ndarray) -> np.ndarray:
    """What ``sf.write(path, y, sr)`` + ``librosa.load(path, sr=sr)`` do to a float32 signal."""
    return pcm16_to_float(float_to_pcm16(y))


def write_wav(path, data: np.ndarray, samplerate: int, subtype: Optional[str] = None) -> None:
    """``soundfile.write`` restricted to what the path uses: mono/multi float data -> PCM_16 WAV
    (the default subtype for ``.wav``)."""
    if subtype not in (None, "PCM_16"):
        raise ValueError("oracle write_wav only implements PCM_16")
    data = np.asarray(data)
    if data.dtype.kind == "f":
        pcm = float_to_pcm16(data)
    elif data.dtype == np.int16:
        pcm = data
    else:
        raise ValueError(f"unsupported dtype {data.dtype}")
    nch = 1 if pcm.ndim == 1 else pcm.shape[1]
    with wave.open(str(path), "wb") as w:
        w.setnchannels(nch)
        w.setsampwidth(2)
        w.setframerate(int(samplerate))
        w.writeframes(np.ascontiguousarray(pcm).astype("<i2").tobytes())


def read_wav(path) -> Tuple[np.ndarray, int]:
    """PCM WAV -> float32 ``[frames, channels]`` the way libsndfile normalises it."""
    with wave.open(str(path), "rb") as w:
        nch, width, sr, nfr = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(nfr)
    if width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) * np.float32(1.0 / 32768.0)
    elif width == 4:
        x = (np.frombuffer(raw, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    elif width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) * np.float32(1.0 / 128.0)
    else:
        raise ValueError(f"unsupported sample width {width}")
    return x.reshape(-1, nch), sr


def load(path, sr: Optional[int] = 22050, mono: bool = True, offset=0.0, duration=None,
         dtype=np.float32, res_type="kaiser_best") -> Tuple[np.ndarray, int]:
    """``librosa.load``: decode -> float32, channel mean if ``mono``; resampling (resampy
    ``kaiser_best``) is only needed when the file rate differs and is outside the hot path
    (all chunks are 48 kHz) -- it raises here."""
    x, sr_native = read_wav(path)
    y = x.T  # [channels, frames]
    if offset:
        y = y[:, int(round(offset * sr_native)):]
    if duration is not None:
        y = y[:, : int(round(duration * sr_native))]
    y = np.mean(y, axis=0) if mono else (y[0] if y.shape[0] == 1 else y)
    if sr is not None and sr != sr_native:
        y = resample(np.ascontiguousarray(y, dtype=dtype), sr_native, sr, res_type=res_type)
    return np.ascontiguousarray(y, dtype=dtype), (sr if sr is not None else sr_native)


# ----------------------------------------------------------------------------------------
# librosa 0.9.2 core/audio.py::resample with res_type="kaiser_best" = resampy 0.2.2 (the versions the reference
# pins: requirements-thesis-baseline-macos-arm64.txt:33 and the resampy that librosa 0.9.2 depends on).  Neither package
# is installed here; this restates their published algorithm:
#   resampy/filters.py::sinc_window   interp_win = taper * rolloff * sinc(rolloff * t), t = linspace(0, num_zeros, n + 1),
#                                     taper = right half of scipy.signal.kaiser(2 n + 1, beta)
#                                     kaiser_best: num_zeros 64, precision 9 (512 table samples per zero crossing),
#                                     rolloff 0.9475937167399596, beta 14.769656459379492
#   resampy/interpn.py::resample_f    per output sample a left and a right filter wing over the input with linear
#                                     interpolation between table entries; y (the input's dtype: float32) is updated in
#                                     place, so every tap's contribution is rounded to float32 as it is added
#   librosa resample                  n_out = ceil(n * ratio), fix_length, no rescaling (scale=False)
# PARITY UNPINNED against resampy itself; cross-checked against torchaudio's Kaiser resampler with the same parameters
# (tests/test_oracle_pinning.py).
# ----------------------------------------------------------------------------------------
KAISER_BEST = dict(num_zeros=64, precision=9, rolloff=0.9475937167399596, beta=14.769656459379492)


def resampy_filter(num_zeros=64, precision=9, rolloff=0.9475937167399596, beta=14.769656459379492):
    """-> (interp_win float64 [num_zeros * 2^precision + 1], num_table = 2^precision)."""
    from scipy.signal.windows import kaiser
    num_bits = 2 ** precision
    n = num_bits * num_zeros
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True))
    taper = kaiser(2 * n + 1, beta)[n:]
    return taper * sinc_win, num_bits


def resample(y: np.ndarray, orig_sr: int, target_sr: int, res_type: str = "kaiser_best") -> np.ndarray:
    """``librosa.resample(y, orig_sr=, target_sr=, res_type="kaiser_best")`` for a 1-D float32 signal."""
    if res_type != "kaiser_best":
        raise NotImplementedError(res_type)
    y = np.ascontiguousarray(y, dtype=np.float32)
    if orig_sr == target_sr:
        return y
    ratio = float(target_sr) / orig_sr
    n_out = int(np.ceil(y.shape[-1] * ratio))
    interp_win, num_table = resampy_filter(**KAISER_BEST)
    interp_win = interp_win.copy()
    if ratio < 1:
        interp_win *= ratio
    interp_delta = np.zeros_like(interp_win)
    interp_delta[:-1] = np.diff(interp_win)
    n_res = int(y.shape[0] * ratio)                    # resampy: shape[axis] = int(shape[axis] * sample_ratio)
    out = np.zeros(n_res, dtype=np.float32)
    scale = min(1.0, ratio)
    time_increment = 1.0 / ratio
    index_step = int(scale * num_table)
    nwin = interp_win.shape[0]
    n_orig = y.shape[0]
    t = np.arange(n_res)
    time_register = np.zeros(n_res)                    # accumulated by repeated addition in the reference loop
    acc = 0.0
    for i in range(n_res):
        time_register[i] = acc
        acc += time_increment
    n = time_register.astype(np.int64)
    frac = scale * (time_register - n)
    for wing in (0, 1):
        if wing == 1:
            frac = scale - frac
        index_frac = frac * num_table
        offset = index_frac.astype(np.int64)
        eta = index_frac - offset
        limit = (nwin - offset) // index_step
        kmax = np.minimum(n + 1, limit) if wing == 0 else np.minimum(n_orig - n - 1, limit)
        for k in range(int(kmax.max()) if kmax.size else 0):
            live = k < kmax
            idx = np.where(live, offset + k * index_step, 0)
            weight = interp_win[idx] + eta * interp_delta[idx]
            src = np.where(live, n - k if wing == 0 else n + k + 1, 0)
            contrib = np.where(live, weight * y[src].astype(np.float64), 0.0)
            out = np.where(live, (out.astype(np.float64) + contrib).astype(np.float32), out)
    # librosa: util.fix_length(y_hat, size=n_out)
    if out.shape[0] < n_out:
        out = np.pad(out, (0, n_out - out.shape[0]))
    return np.ascontiguousarray(out[:n_out], dtype=np.float32)
