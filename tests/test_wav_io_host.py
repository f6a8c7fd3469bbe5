"""CPU: WAV decoding at the edge of the path (row N2) -- ``reference_api.load_wav`` = ``librosa.load(sr=sr, mono=True)`` on
WAV files: libsndfile's integer -> float conversions (PCM_16 / 2^15, PCM_24 / 2^23, PCM_32 / 2^31, PCM_U8 (u - 128) / 2^7),
IEEE float as stored, channel mean, and the PCM_16 writer's ``lrintf(x * 0x7FFF)`` (00:55-57)."""
import struct
import wave

import numpy as np
import pytest

from amphibian_vae_latent_detector_b200 import reference_api as api


def _pcm(path, frames: bytes, width, nch=1, sr=48000):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(nch)
        w.setsampwidth(width)
        w.setframerate(sr)
        w.writeframes(frames)


def _float_wav(path, x, bits=32, extensible=False, nch=1, sr=48000):
    data = x.astype("<f4" if bits == 32 else "<f8").tobytes()
    core = struct.pack("<HHIIHH", 0xFFFE if extensible else 3, nch, sr, sr * nch * bits // 8, nch * bits // 8, bits)
    if extensible:
        core += struct.pack("<HHI", 22, bits, 0) + struct.pack("<H", 3) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(core)) + core + b"fact" + struct.pack("<II", 4, x.shape[0] // nch) + \
        b"data" + struct.pack("<I", len(data)) + data
    path.write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)


def test_pcm16_is_value_over_32768(tmp_path):
    v = np.array([-32768, -32767, -1, 0, 1, 12345, 32767], dtype="<i2")
    _pcm(tmp_path / "a.wav", v.tobytes(), 2)
    assert np.array_equal(api.load_wav(tmp_path / "a.wav"), v.astype(np.float32) / np.float32(32768.0))


def test_pcm24_pcm32_u8(tmp_path):
    v24 = np.array([-8388608, -8388607, -65537, -1, 0, 1, 255, 65536, 8388607], dtype=np.int64)
    raw = b"".join(int(x & 0xFFFFFF).to_bytes(3, "little") for x in v24)
    _pcm(tmp_path / "b.wav", raw, 3)
    got = api.load_wav(tmp_path / "b.wav")
    assert got.dtype == np.float32 and np.array_equal(got.astype(np.float64), v24 / 8388608.0)        # exact in float32
    v32 = np.array([-2147483648, -2147483647, -1, 0, 1, 16777217, 2147483647], dtype="<i4")
    _pcm(tmp_path / "c.wav", v32.tobytes(), 4)
    assert np.array_equal(api.load_wav(tmp_path / "c.wav"), (v32.astype(np.float32) * np.float32(2.0 ** -31)))
    u8 = np.array([0, 1, 127, 128, 129, 255], dtype=np.uint8)
    _pcm(tmp_path / "d.wav", u8.tobytes(), 1)
    assert np.array_equal(api.load_wav(tmp_path / "d.wav"), (u8.astype(np.float32) - 128) / 128)


@pytest.mark.parametrize("bits,ext", [(32, False), (32, True), (64, False)])
def test_ieee_float_as_stored(tmp_path, bits, ext):
    x = np.random.default_rng(3).standard_normal(1001).astype(np.float32) * np.float32(0.3)
    _float_wav(tmp_path / "f.wav", x, bits=bits, extensible=ext)
    assert np.array_equal(api.load_wav(tmp_path / "f.wav"), x)


def test_stereo_is_averaged(tmp_path):
    v = np.array([[100, 300], [-32768, 32767], [5, 6]], dtype="<i2")
    _pcm(tmp_path / "s.wav", v.tobytes(), 2, nch=2)
    f = v.astype(np.float32) / np.float32(32768.0)
    assert np.array_equal(api.load_wav(tmp_path / "s.wav"), np.mean(f.T, axis=0))                      # librosa.to_mono
    (tmp_path / "junk.wav").write_bytes(b"RIFFxxxxWAVEjunk")
    with pytest.raises(Exception):
        api.load_wav(tmp_path / "junk.wav")


def test_writer_round_trip_is_lrint_of_x_times_7fff(tmp_path):
    x = np.array([-1.0, -0.99998, -0.5, -1.5e-5, 0.0, 1.5e-5, 4.5 / 32767, 0.5, 0.99998, 1.0], dtype=np.float32)
    api.write_wav_pcm16(tmp_path / "w.wav", x, 48000)
    with wave.open(str(tmp_path / "w.wav"), "rb") as w:
        q = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
    assert np.array_equal(q, np.rint(x * np.float32(32767.0)).astype(np.int16))                        # round half to even
    assert np.array_equal(api.load_wav(tmp_path / "w.wav"), q.astype(np.float32) / np.float32(32768.0))
