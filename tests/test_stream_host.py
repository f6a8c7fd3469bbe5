"""CPU: windowing of a long recording -- WAV payload mapping, window starts, right
zero-padding of the last partial window (map_detector_core.py:214-215), overlapping hops, slab double buffering."""
import struct
import wave

import numpy as np
import pytest

from amphibian_vae_latent_detector_b200 import stream


def _write(path, x, sr=48000, nch=1, width=2):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(nch)
        w.setsampwidth(width)
        w.setframerate(sr)
        w.writeframes(x.tobytes())


def test_open_maps_payload_exactly(tmp_path):
    x = np.random.default_rng(0).integers(-32768, 32767, 123457).astype("<i2")
    _write(tmp_path / "a.wav", x)
    m = stream.open_pcm16_mono(tmp_path / "a.wav", 48000)
    assert m.dtype == np.dtype("<i2") and np.array_equal(m, x)


def test_open_skips_extra_chunks(tmp_path):
    x = np.arange(-50, 50).astype("<i2")
    body = b"WAVE" + b"LIST" + struct.pack("<I", 5) + b"abcde\x00" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, 48000, 96000, 2, 16) \
        + b"data" + struct.pack("<I", x.nbytes) + x.tobytes()
    (tmp_path / "b.wav").write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
    assert np.array_equal(stream.open_pcm16_mono(tmp_path / "b.wav", 48000), x)


def test_open_rejects_other_formats(tmp_path):
    _write(tmp_path / "st.wav", np.zeros(200, "<i2"), nch=2)
    with pytest.raises(ValueError):
        stream.open_pcm16_mono(tmp_path / "st.wav", 48000)
    _write(tmp_path / "sr.wav", np.zeros(200, "<i2"), sr=44100)
    with pytest.raises(RuntimeError):
        stream.open_pcm16_mono(tmp_path / "sr.wav", 48000)
    (tmp_path / "junk.wav").write_bytes(b"not a wav file at all")
    with pytest.raises(RuntimeError):
        stream.open_pcm16_mono(tmp_path / "junk.wav", 48000)


def test_window_starts():
    assert stream.window_starts(0, 10, 10).tolist() == []
    assert stream.window_starts(10, 10, 10).tolist() == [0]
    assert stream.window_starts(11, 10, 10).tolist() == [0, 10]
    assert stream.window_starts(25, 10, 5).tolist() == [0, 5, 10, 15, 20]


@pytest.mark.parametrize("hop", [3000, 1000, 4500])
@pytest.mark.parametrize("slab", [1, 2, 7, 100])
def test_slabs_cover_the_recording(hop, slab):
    x = (np.arange(10007) % 30000 - 15000).astype(np.int16)
    L = 3000
    seen = []
    for starts, buf in stream.iter_slabs(x, L, hop, slab):
        assert buf.dtype.is_floating_point is False and buf.shape == (len(starts), L)
        for s, row in zip(starts, buf.numpy()):
            m = min(L, x.shape[0] - int(s))
            assert np.array_equal(row[:m], x[s:s + m]) and not row[m:].any()
            seen.append(int(s))
    assert seen == stream.window_starts(x.shape[0], L, hop).tolist()
