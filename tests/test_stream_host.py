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


def _rf64(path, n_samples, tail, *, magic=b"RF64", data_field=0xFFFFFFFF, with_ds64=True):
    """A sparse 64-bit WAV: header, ``n_samples`` zero samples (a hole in the file), the last ``len(tail)`` of them = tail."""
    data_bytes = 2 * n_samples
    ds64 = b"ds64" + struct.pack("<IQQQI", 28, 4 + 36 + 24 + data_bytes, data_bytes, n_samples, 0) if with_ds64 else b""
    head = magic + struct.pack("<I", 0xFFFFFFFF) + b"WAVE" + ds64 + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, 48000, 96000, 2, 16) \
        + b"data" + struct.pack("<I", data_field)
    with open(path, "wb") as f:
        f.write(head)
        f.truncate(len(head) + data_bytes)
        f.seek(len(head) + data_bytes - tail.nbytes)
        f.write(tail.tobytes())
    return len(head)


@pytest.mark.parametrize("kind", ["rf64", "bw64", "riff_placeholder", "riff_zero"])
def test_open_day_long_recordings_beyond_4_gib(tmp_path, kind):
    """24 h at 48 kHz = 4 147 200 000 samples = 8.3 GB of PCM_16: RIFF's 32-bit sizes cannot hold it (BASELINE configs[4])."""
    n = 24 * 3600 * 48000
    tail = np.arange(-500, 500).astype("<i2")
    path = tmp_path / f"{kind}.wav"
    kw = {"rf64": {}, "bw64": {"magic": b"BW64"}, "riff_placeholder": {"magic": b"RIFF", "with_ds64": False},
          "riff_zero": {"magic": b"RIFF", "with_ds64": False, "data_field": 0}}[kind]
    _rf64(path, n, tail, **kw)
    if path.stat().st_blocks * 512 > 64 << 20:
        pytest.skip("file system without sparse files")
    m = stream.open_pcm16_mono(path, 48000)
    assert m.shape == (n,)
    assert np.array_equal(m[-1000:], tail) and not m[:4096].any() and not m[n // 2:n // 2 + 4096].any()
    starts = stream.window_starts(n, 144000, 144000)
    assert starts.shape == (28800,) and int(starts[-1]) + 144000 == n              # 28 800 windows of 3 s


def test_open_rf64_data_size_shorter_than_file(tmp_path):
    """ds64 holds the true payload size: trailing chunks after the data are not read as samples."""
    x = np.arange(-100, 100).astype("<i2")
    ds64 = b"ds64" + struct.pack("<IQQQI", 28, 0, x.nbytes, x.shape[0], 0)
    body = b"WAVE" + ds64 + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, 48000, 96000, 2, 16) + b"data" + \
        struct.pack("<I", 0xFFFFFFFF) + x.tobytes() + b"LIST" + struct.pack("<I", 4) + b"tail"
    (tmp_path / "c.wav").write_bytes(b"RF64" + struct.pack("<I", 0xFFFFFFFF) + body)
    assert np.array_equal(stream.open_pcm16_mono(tmp_path / "c.wav", 48000), x)
