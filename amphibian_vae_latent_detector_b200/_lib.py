"""ctypes binding of ``libavld.so`` (the C ABI declared in ``include/avld.h``).

The library is built in-tree by ``amphibian_vae_latent_detector_b200.build`` and loaded from the
package directory.  There is no fallback of any kind: if the shared object is missing, or a call
fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

HERE = Path(__file__).resolve().parent
# AVLD_LIB_PATH selects another build of the same library (tools/ point it at libavld_bringup.so); it is never a fallback
LIB_PATH = Path(os.environ["AVLD_LIB_PATH"]).resolve() if os.environ.get("AVLD_LIB_PATH") else HERE / "libavld.so"

OK = 0
ERROR_NAMES = {-1: "AVLD_ERR_INVALID", -2: "AVLD_ERR_CUDA", -3: "AVLD_ERR_UNSUPPORTED", -4: "AVLD_ERR_STATE",
               -5: "AVLD_ERR_NOMEM"}


class AvldError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("sr", C.c_int32), ("chunk_len", C.c_int32), ("n_fft", C.c_int32), ("hop", C.c_int32),
                ("n_mels", C.c_int32), ("fmin", C.c_float), ("fmax", C.c_float), ("target_frames", C.c_int32),
                ("amin", C.c_float), ("top_db", C.c_float), ("max_batch", C.c_int32)]


class Layer(C.Structure):
    _fields_ = [("kind", C.c_int32), ("c_in", C.c_int32), ("c_out", C.c_int32), ("ksize", C.c_int32),
                ("stride", C.c_int32), ("pad", C.c_int32), ("relu", C.c_int32), ("pool", C.c_int32),
                ("in_h", C.c_int32), ("in_w", C.c_int32), ("weight", C.POINTER(C.c_float)),
                ("bias", C.POINTER(C.c_float))]


class Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("in0", C.c_int32), ("in1", C.c_int32), ("out", C.c_int32), ("c_in", C.c_int32),
                ("c_out", C.c_int32), ("ksize", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32), ("relu", C.c_int32),
                ("pool", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32), ("weight", C.POINTER(C.c_float)),
                ("bias", C.POINTER(C.c_float))]


OP_CONV, OP_LINEAR, OP_ADD, OP_AFFINE, OP_POOL = range(5)
COMM_ID_BYTES = 128


class RankQuery(C.Structure):
    _fields_ = [("species", C.c_int32), ("side", C.c_int32), ("rank", C.c_int64)]


_P = C.c_void_p
_SIGNATURES = {
    "avld_abi_version": (C.c_int, []),
    "avld_last_error": (C.c_char_p, []),
    "avld_ctx_create": (C.c_int, [C.c_int, C.POINTER(Params), C.POINTER(_P)]),
    "avld_ctx_destroy": (None, [_P]),
    "avld_ctx_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "avld_ctx_set_normalization": (C.c_int, [_P, C.c_int, C.c_double, C.c_double, C.c_double]),
    "avld_ctx_dft_info": (C.c_int, [_P, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "avld_profile_enable": (C.c_int, [_P, C.c_int]),
    "avld_profile_collect": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_uint64), C.c_int]),
    "avld_stage_count": (C.c_int, []),
    "avld_stage_name": (C.c_char_p, [C.c_int]),
    "avld_rms_normalize": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int, _P]),
    "avld_resample_len": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32]),
    "avld_resample": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "avld_logmel": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "avld_normalize_logmel": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int, _P]),
    "avld_encoder_load": (C.c_int, [_P, C.POINTER(Layer), C.c_int32]),
    "avld_encoder_load_program": (C.c_int, [_P, C.POINTER(Op), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "avld_encoder_forward": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "avld_encode": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int, _P]),
    "avld_encode_pcm16": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int, _P]),
    "avld_centroid_accumulate": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "avld_radii": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "avld_order_stats": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.POINTER(RankQuery), C.c_int32,
                                   C.POINTER(C.c_float), _P]),
    "avld_comm_unique_id": (C.c_int, [_P]),
    "avld_comm_init": (C.c_int, [_P, _P, C.c_int32, C.c_int32]),
    "avld_comm_destroy": (C.c_int, [_P]),
    "avld_allreduce_centroids": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "avld_allgather_radii": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int32, _P, _P, _P]),
    "avld_decide": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, _P]),
    "avld_map_score": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_double, C.c_int, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "avld_cov_accumulate": (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "avld_encode_detect_host": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, _P, _P, C.c_int32, _P, _P, _P, _P]),
    "avld_encode_detect_host_pcm16": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, _P, _P, C.c_int32, _P, _P, _P, _P]),
    "avld_pairwise_plan": (C.c_int64, [C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int64]),
    "avld_mel_taps": (C.c_int, [C.POINTER(Params), C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "avld_dbg_gemm": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load (once) and type the shared library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m amphibian_vae_latent_detector_b200.build` "
                          "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI and the header ever diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(code: int) -> None:
    if code != OK:
        raise AvldError(code, load().avld_last_error().decode("utf-8", "replace"))
