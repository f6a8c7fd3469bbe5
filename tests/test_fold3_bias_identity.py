"""CPU check of the arithmetic identity fold3_kernel<2> relies on (csrc/fold3.cu, Samples<2>): the normalised PCM_16 integers
are handed out as the float32 2^23 + 2^15 + q ("biased", one PRMT from the stored uint16 q + 32768) and every fold term is
formed as a - b = biased(a) - biased(b) and a + b = (biased(a) - 2 * bias) + biased(b).  Both must be exact in float32 for
every pair of 16-bit samples, i.e. equal to the integer sum / difference the scalar form (un-bias first) computes."""
import numpy as np

BIAS = np.float32(8421376.0)      # 2^23 + 2^15


def _biased(q):
    u = (q.astype(np.int32) + 32768).astype(np.uint32)
    return (np.uint32(0x4B000000) | u).view(np.float32)        # the float 2^23 + u, as the kernel's PRMT builds it


def test_bias_constants_are_exact_floats():
    assert float(BIAS) == 2.0 ** 23 + 2.0 ** 15
    assert float(np.float32(2.0) * BIAS) == 2.0 ** 24 + 2.0 ** 16
    q = np.array([-32768, -1, 0, 1, 32767], dtype=np.int16)
    assert np.array_equal(_biased(q) - BIAS, q.astype(np.float32))


def test_biased_sum_and_difference_are_exact():
    rng = np.random.default_rng(7)
    edge = np.array([-32768, -32767, -1, 0, 1, 32766, 32767], dtype=np.int16)
    a = np.concatenate([np.repeat(edge, edge.size), rng.integers(-32768, 32768, 200000).astype(np.int16)])
    b = np.concatenate([np.tile(edge, edge.size), rng.integers(-32768, 32768, 200000).astype(np.int16)])
    ma, mb = _biased(a), _biased(b)
    two_bias = np.float32(2.0) * BIAS
    s = (ma - two_bias) + mb              # float32 arithmetic, one rounding per operation
    d = ma - mb
    assert s.dtype == np.float32 and d.dtype == np.float32
    assert np.array_equal(s, (a.astype(np.int32) + b.astype(np.int32)).astype(np.float32))
    assert np.array_equal(d, (a.astype(np.int32) - b.astype(np.int32)).astype(np.float32))
    # a "zero" sample in biased form (the taps that pair with nothing, out-of-range padding) is the bias itself
    z = np.full_like(ma, BIAS)
    assert np.array_equal((ma - two_bias) + z, a.astype(np.float32))
    assert np.array_equal((z - two_bias) + mb, b.astype(np.float32))
    assert np.array_equal(ma - z, a.astype(np.float32))
