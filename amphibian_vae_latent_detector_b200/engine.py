"""Host-side engine: torch tensors in, torch tensors out, every operation a call into ``libavld.so``.

PyTorch is used for device memory, streams and ``torch.distributed`` only.  There is no fallback: a
missing library raises at import of :mod:`_lib`, a failing call raises :class:`AvldError`, CPU tensors
are rejected (except by :meth:`Engine.encode_detect_host`, whose contract is host buffers).

Reference functions behind each method (paths relative to ``latent_space_exploration/``):

=============================  =====================================================================
``rms_normalize``              ``rms_normalize`` 00_normalize_dataset_rms.py:29-38 (+ ``sf.write`` :57)
``logmel`` / ``normalize_logmel``  ``wav_to_mel`` map_detector_core.py:198-237 after the file load
``load_encoder`` / ``encoder_forward``  ``load_encoder`` core:150-179, ``encoder(x)`` core:270-300
``encode``                     ``encode_wav_to_latent`` core:240-300 on in-memory chunks
``centroid_accumulate``/``radii``/``order_stats``  ``fit_species_with_fp_control`` 08:310-333
``decide``                     ``detect_species`` 09:416-436, ``predict_one`` 10:175-199
``fit_radial``                 the fit loop 08:530-558 (+ the q_out grid of run_qout_grid.sh:13)
=============================  =====================================================================
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .radial_fit import RadialFit, fit_radial
from .map_fit import MapFit, fit_map
from .encoder import AddOp, AffineOp, ConvOp, EncoderProgram, LinearOp, PoolOp, export_program

DEFAULTS = dict(sr=48000, n_fft=2048, hop_length=384, n_mels=64, fmin=150.0, fmax=15000.0, target_frames=192,
                amin=1e-10, top_db=80.0)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class Engine:
    def __init__(self, device: int | torch.device = 0, *, chunk_len: int = 144000, max_batch: int = 256,
                 sr: int = 48000, n_fft: int = 2048, hop_length: int = 384, n_mels: int = 64, fmin: float = 150.0,
                 fmax: float = 15000.0, target_frames: int = 192, amin: float = 1e-10, top_db: float = 80.0,
                 scalar_semantics: str = "numpy2", target_rms: float = 0.05, rms_min: float = 1e-4, eps: float = 1e-8):
        """``scalar_semantics``: how ``rms + eps`` and ``target_rms / (...)`` of rms_normalize (00:29-38) are rounded --
        "numpy2" (float32, what numpy >= 2 does and the committed fixtures were made with) or "numpy1" (float64, what the
        reference's pinned numpy==1.26.4 does).  ``target_rms / rms_min / eps`` are the constants of the fused host calls."""
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("amphibian_vae_latent_detector_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device("cuda", device if isinstance(device, int) else (device.index or 0))
        self.params = _lib.Params(sr, chunk_len, n_fft, hop_length, n_mels, fmin, fmax, target_frames, amin, top_db,
                                  max_batch)
        self.chunk_len, self.max_batch = int(chunk_len), int(max_batch)
        self.target_frames, self.n_mels = int(target_frames), int(n_mels)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.avld_ctx_create(self.device.index, C.byref(self.params), C.byref(h)))
        self._h = h
        if scalar_semantics not in ("numpy1", "numpy2"):
            raise ValueError("scalar_semantics must be 'numpy1' or 'numpy2'")
        self.scalar_semantics = scalar_semantics
        self.norm = (float(target_rms), float(rms_min), float(eps))
        _lib.check(self.lib.avld_ctx_set_normalization(self._h, 1 if scalar_semantics == "numpy1" else 0, *self.norm))
        nf, ld, sm = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(self.lib.avld_ctx_info(self._h, C.byref(nf), C.byref(ld), C.byref(sm)))
        self.n_frames, self.sm_count = nf.value, sm.value
        self.latent_dim = 0
        self.program: Optional[EncoderProgram] = None

    # ------------------------------------------------------------------ life cycle
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.avld_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _dev(self, t: torch.Tensor, dtype: torch.dtype, name: str) -> torch.Tensor:
        if not isinstance(t, torch.Tensor) or t.device != self.device:
            raise ValueError(f"{name} must be a tensor on {self.device} (got {getattr(t, 'device', type(t))})")
        if t.dtype != dtype:
            raise ValueError(f"{name} must be {dtype} (got {t.dtype})")
        return t.contiguous()

    # ------------------------------------------------------------------ accounting
    def profile(self, on: bool) -> None:
        _lib.check(self.lib.avld_profile_enable(self._h, int(on)))

    def collect(self, reset: bool = True) -> Dict[str, Dict[str, float]]:
        """-> ``{kernel family: {ms, timed_launches, launches}}`` (launch counts since the last reset)."""
        ns = self.lib.avld_stage_count()
        ms = (C.c_double * ns)()
        tl = (C.c_int64 * ns)()
        ln = (C.c_uint64 * ns)()
        _lib.check(self.lib.avld_profile_collect(self._h, ms, tl, ln, int(reset)))
        return {self.lib.avld_stage_name(i).decode(): dict(ms=ms[i], timed_launches=int(tl[i]), launches=int(ln[i]))
                for i in range(ns)}

    def dft_info(self) -> Dict[str, object]:
        """How the STFT is evaluated: ``{mode, algorithmic_flops_per_chunk, issued_flops_per_chunk}``."""
        mode = C.c_char_p()
        alg, iss = C.c_double(), C.c_double()
        _lib.check(self.lib.avld_ctx_dft_info(self._h, C.byref(mode), C.byref(alg), C.byref(iss)))
        return {"mode": mode.value.decode(), "algorithmic_flops_per_chunk": alg.value, "issued_flops_per_chunk": iss.value}

    def _norm_args(self, target_rms, rms_min, eps):
        t, r, e = self.norm
        return (t if target_rms is None else float(target_rms), r if rms_min is None else float(rms_min),
                e if eps is None else float(eps))

    # ------------------------------------------------------------------ R1 / R2
    def rms_normalize(self, x: torch.Tensor, target_rms: Optional[float] = None, rms_min: Optional[float] = None,
                      eps: Optional[float] = None,
                      pcm16: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """-> ``(y [n,L] f32, ok [n] uint8, rms [n] f32)``; bit-exact with numpy's float32 arithmetic under the engine's
        ``scalar_semantics``.  Constants default to the engine's (``Engine(target_rms=..., rms_min=..., eps=...)``)."""
        target_rms, rms_min, eps = self._norm_args(target_rms, rms_min, eps)
        x = self._dev(x, torch.float32, "x")
        if x.ndim != 2 or x.shape[1] != self.chunk_len:
            raise ValueError(f"x must be [n, {self.chunk_len}]")
        n = x.shape[0]
        y = torch.empty_like(x)
        ok = torch.empty(n, dtype=torch.uint8, device=self.device)
        rms = torch.empty(n, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.avld_rms_normalize(self._h, _ptr(x), _ptr(y), _ptr(ok), _ptr(rms), n, target_rms, rms_min,
                                               eps, int(pcm16), _stream()))
        return y, ok, rms

    # ------------------------------------------------------------------ M2-M5 + E0
    def resample(self, y: torch.Tensor, sr_in: int, sr_out: int) -> torch.Tensor:
        """``librosa.resample(y, orig_sr=sr_in, target_sr=sr_out)`` (``kaiser_best``) of a mono float32 signal ``[n]``."""
        y = self._dev(y, torch.float32, "y")
        if y.ndim != 1:
            raise ValueError("y must be a 1-D mono signal")
        n_out = int(self.lib.avld_resample_len(y.shape[0], int(sr_in), int(sr_out)))
        out = torch.empty(n_out, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.avld_resample(self._h, _ptr(y), y.shape[0], int(sr_in), int(sr_out), _ptr(out), n_out, _stream()))
        return out

    def logmel(self, y: torch.Tensor) -> torch.Tensor:
        """-> features ``[n, T, M]`` float32 (= ``wav_to_mel(...).T`` per chunk)."""
        y = self._dev(y, torch.float32, "y")
        if y.ndim != 2 or y.shape[1] != self.chunk_len:
            raise ValueError(f"y must be [n, {self.chunk_len}]")
        feat = torch.empty(y.shape[0], self.target_frames, self.n_mels, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.avld_logmel(self._h, _ptr(y), _ptr(feat), y.shape[0], _stream()))
        return feat

    def normalize_logmel(self, x: torch.Tensor, target_rms: Optional[float] = None, rms_min: Optional[float] = None,
                         eps: Optional[float] = None,
                         pcm16: bool = True) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        target_rms, rms_min, eps = self._norm_args(target_rms, rms_min, eps)
        x = self._dev(x, torch.float32, "x")
        n = x.shape[0]
        feat = torch.empty(n, self.target_frames, self.n_mels, dtype=torch.float32, device=self.device)
        ok = torch.empty(n, dtype=torch.uint8, device=self.device)
        rms = torch.empty(n, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.avld_normalize_logmel(self._h, _ptr(x), _ptr(feat), _ptr(ok), _ptr(rms), n, target_rms,
                                                  rms_min, eps, int(pcm16), _stream()))
        return feat, ok, rms

    # ------------------------------------------------------------------ L1 / E1 / E2
    def load_encoder(self, module_or_program) -> EncoderProgram:
        """Accepts the ``nn.Module`` the reference's ``load_encoder`` returns (core:150-179) or an
        already exported :class:`EncoderProgram`."""
        prog = module_or_program if isinstance(module_or_program, EncoderProgram) else \
            export_program(module_or_program, self.target_frames, self.n_mels)
        if prog.in_hw[0] * prog.n_seg != self.target_frames or prog.in_hw[1] != self.n_mels:
            raise ValueError(f"the program was exported for {prog.n_seg} x {prog.in_hw} feature segments, the engine makes "
                             f"[{self.target_frames}, {self.n_mels}] images")
        ops = (_lib.Op * len(prog.ops))()
        keep = []

        def fptr(a):
            a = np.ascontiguousarray(a, dtype=np.float32)
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_float))

        null = C.POINTER(C.c_float)()
        for i, op in enumerate(prog.ops):
            o = ops[i]
            o.in1, o.ksize, o.stride, o.pad, o.relu, o.pool, o.in_h, o.in_w = -1, 1, 1, 0, 0, 0, 1, 1
            o.weight, o.bias = null, null
            if isinstance(op, ConvOp):
                cout, kh, _, cin = op.weight.shape
                o.kind, o.in0, o.out, o.c_in, o.c_out = _lib.OP_CONV, op.src, op.dst, cin, cout
                o.in1 = op.residual                    # fused residual add (-1: none)
                o.ksize, o.stride, o.pad, o.relu = kh, op.stride, op.pad, int(op.relu)
                o.pool = 0 if op.pool == 1 else (2 if op.pool_avg else 1)
                o.in_h, o.in_w = op.in_hw
                o.weight, o.bias = fptr(op.weight), fptr(op.bias)
            elif isinstance(op, LinearOp):
                o.kind, o.in0, o.out = _lib.OP_LINEAR, op.src, op.dst
                o.c_in, o.c_out, o.relu = op.weight.shape[1], op.weight.shape[0], int(op.relu)
                o.weight, o.bias = fptr(op.weight), fptr(op.bias)
            elif isinstance(op, AddOp):
                o.kind, o.in0, o.in1, o.out, o.relu = _lib.OP_ADD, op.a, op.b, op.dst, int(op.relu)
            elif isinstance(op, AffineOp):
                o.kind, o.in0, o.out, o.relu = _lib.OP_AFFINE, op.src, op.dst, int(op.relu)
                o.c_in = o.c_out = int(op.scale.shape[0])
                o.weight, o.bias = fptr(op.scale), fptr(op.shift)
            elif isinstance(op, PoolOp):
                o.kind, o.in0, o.out = _lib.OP_POOL, op.src, op.dst
                o.ksize, o.stride, o.pool = op.k, op.stride, 2 if op.avg else 1
                o.in_h, o.in_w = op.in_hw
            else:
                raise TypeError(type(op))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.avld_encoder_load_program(self._h, ops, len(prog.ops), prog.out, int(prog.out_nchw is not None),
                                                          prog.in_hw[0], prog.n_seg))
        self.latent_dim = prog.latent_dim
        self.program = prog
        return prog

    def encoder_forward(self, feat: torch.Tensor) -> torch.Tensor:
        feat = self._dev(feat, torch.float32, "feat")
        if feat.ndim == 4 and feat.shape[1] == 1:
            feat = feat[:, 0]
        if feat.ndim != 3 or tuple(feat.shape[1:]) != (self.target_frames, self.n_mels):
            raise ValueError(f"feat must be [n, {self.target_frames}, {self.n_mels}]")
        feat = feat.contiguous()
        mu = torch.empty(feat.shape[0], self.latent_dim, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.avld_encoder_forward(self._h, _ptr(feat), _ptr(mu), feat.shape[0], _stream()))
        return mu

    def encode(self, x: torch.Tensor, *, pcm16: bool = True, target_rms: Optional[float] = None,
               rms_min: Optional[float] = None, eps: Optional[float] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """raw chunks ``[n, L]`` (float32, or int16 PCM_16 samples as the WAV files hold them) -> ``(mu [n, D], ok [n])``:
        normalise (+ PCM_16 round trip of the ``*_norm`` dataset on disk) -> log-mel -> encoder."""
        target_rms, rms_min, eps = self._norm_args(target_rms, rms_min, eps)
        as_pcm = isinstance(x, torch.Tensor) and x.dtype == torch.int16
        x = self._dev(x, torch.int16 if as_pcm else torch.float32, "x")
        n = x.shape[0]
        mu = torch.empty(n, self.latent_dim, dtype=torch.float32, device=self.device)
        ok = torch.empty(n, dtype=torch.uint8, device=self.device)
        fn = self.lib.avld_encode_pcm16 if as_pcm else self.lib.avld_encode
        _lib.check(fn(self._h, _ptr(x), _ptr(mu), _ptr(ok), n, target_rms, rms_min, eps, int(pcm16), _stream()))
        return mu, ok

    # ------------------------------------------------------------------ F1-F3
    def centroid_accumulate(self, Z: torch.Tensor, label: torch.Tensor, K: int,
                            out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        Z = self._dev(Z, torch.float32, "Z")
        label = self._dev(label, torch.int32, "label")
        D = Z.shape[1]
        if out is None:
            out = (torch.zeros(K, D, dtype=torch.float64, device=self.device),
                   torch.zeros(K, dtype=torch.int64, device=self.device))
        _lib.check(self.lib.avld_centroid_accumulate(self._h, _ptr(Z), _ptr(label), _ptr(out[0]), _ptr(out[1]),
                                                     Z.shape[0], K, D, _stream()))
        return out

    def radii(self, Z: torch.Tensor, centroid: torch.Tensor) -> torch.Tensor:
        Z = self._dev(Z, torch.float32, "Z")
        centroid = self._dev(centroid, torch.float32, "centroid")
        K, D = centroid.shape
        if Z.shape[1] != D:
            raise ValueError("latent / centroid dimension mismatch")
        r = torch.empty(Z.shape[0], K, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.avld_radii(self._h, _ptr(Z), _ptr(centroid), _ptr(r), Z.shape[0], K, D, _stream()))
        return r

    def order_stats(self, radii: torch.Tensor, label: torch.Tensor,
                    queries: Sequence[Tuple[int, int, int]]) -> np.ndarray:
        """``queries`` = (species, side, rank) triples -> float32 values (exact selection)."""
        radii = self._dev(radii, torch.float32, "radii")
        label = self._dev(label, torch.int32, "label")
        arr = (_lib.RankQuery * len(queries))(*[_lib.RankQuery(int(k), int(s), int(r)) for k, s, r in queries])
        out = np.empty(len(queries), dtype=np.float32)
        _lib.check(self.lib.avld_order_stats(self._h, _ptr(radii), _ptr(label), radii.shape[0], radii.shape[1], arr,
                                             len(queries), out.ctypes.data_as(C.POINTER(C.c_float)), _stream()))
        return out

    # ------------------------------------------------------------------ D2
    # ------------------------------------------------------------------ multi-GPU fit without torch.distributed
    comm_world = 0          # > 1 once comm_init has run: fit_radial(group="avld") then uses the context's own communicator

    def comm_unique_id(self) -> bytes:
        """Rank 0: a fresh NCCL id (128 bytes) to hand to the other ranks by any means (file, socket, queue)."""
        buf = C.create_string_buffer(_lib.COMM_ID_BYTES)
        _lib.check(self.lib.avld_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int) -> None:
        """Collective over all ranks: this context's NCCL communicator (one process and one context per GPU)."""
        if len(unique_id) != _lib.COMM_ID_BYTES:
            raise ValueError(f"the id has {len(unique_id)} bytes, not {_lib.COMM_ID_BYTES}")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.avld_comm_init(self._h, C.c_char_p(unique_id), int(rank), int(world)))
        self.comm_rank, self.comm_world = int(rank), int(world)

    def allreduce_centroids(self, sums: torch.Tensor, cnts: torch.Tensor) -> None:
        """In place: per-species sums [K, D] float64 and counts [K] int64 summed over the ranks."""
        sums = self._dev(sums, torch.float64, "sums")
        cnts = self._dev(cnts, torch.int64, "cnts")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.avld_allreduce_centroids(self._h, _ptr(sums), _ptr(cnts), sums.shape[0], sums.shape[1], _stream()))

    def allgather_radii(self, radii: torch.Tensor, label: torch.Tensor, shard_rows: int):
        """-> (radii_all [world * shard_rows, K], label_all [world * shard_rows]); padding rows carry label -1."""
        radii = self._dev(radii, torch.float32, "radii")
        label = self._dev(label, torch.int32, "label")
        n, K = radii.shape
        ra = torch.empty(self.comm_world * shard_rows, K, dtype=torch.float32, device=self.device)
        la = torch.empty(self.comm_world * shard_rows, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.avld_allgather_radii(self._h, _ptr(radii), _ptr(label), n, int(shard_rows), K, _ptr(ra), _ptr(la), _stream()))
        return ra, la

    def decide(self, radii: torch.Tensor, thr: torch.Tensor, priority_rank: torch.Tensor):
        radii = self._dev(radii, torch.float32, "radii")
        thr = self._dev(thr, torch.float64, "thr")
        priority_rank = self._dev(priority_rank, torch.int32, "priority_rank")
        n, K = radii.shape
        pred = torch.empty(n, dtype=torch.int32, device=self.device)
        best = torch.empty(n, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.avld_decide(self._h, _ptr(radii), _ptr(thr), _ptr(priority_rank), _ptr(pred), _ptr(best), n,
                                        K, _stream()))
        return pred, best

    def encode_detect_host(self, x_host, centroid: np.ndarray, thr: np.ndarray, priority_rank: np.ndarray, *,
                           pcm16: bool = True, want_mu: bool = False):
        """HOST buffers in / out (the call the drop-in layer makes per batch of decoded files):
        ``x_host [n, L]`` float32, or int16 PCM_16 samples as stored in the WAV files (half the PCIe bytes;
        decoded on the GPU as s / 32768), numpy or CPU tensor, pinned for full copy/compute overlap ->
        ``(pred [n] int32, best_d [n] f32, ok [n] uint8, mu [n, D] f32 | None)`` as numpy arrays."""
        xt = x_host if isinstance(x_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x_host))
        if xt.device.type != "cpu" or xt.dtype not in (torch.float32, torch.int16) or xt.ndim != 2 \
                or xt.shape[1] != self.chunk_len:
            raise ValueError(f"x_host must be a CPU float32 or int16 (PCM_16) [n, {self.chunk_len}] array")
        fn = self.lib.avld_encode_detect_host if xt.dtype == torch.float32 else self.lib.avld_encode_detect_host_pcm16
        xt = xt.contiguous()
        n = xt.shape[0]
        centroid = np.ascontiguousarray(centroid, dtype=np.float32)
        thr = np.ascontiguousarray(thr, dtype=np.float64)
        priority_rank = np.ascontiguousarray(priority_rank, dtype=np.int32)
        K = centroid.shape[0]
        pred = np.empty(n, dtype=np.int32)
        best = np.empty(n, dtype=np.float32)
        ok = np.empty(n, dtype=np.uint8)
        mu = np.empty((n, self.latent_dim), dtype=np.float32) if want_mu else None
        _lib.check(fn(
            self._h, xt.data_ptr(), n, int(pcm16), centroid.ctypes.data, thr.ctypes.data, priority_rank.ctypes.data, K,
            pred.ctypes.data, best.ctypes.data, None if mu is None else mu.ctypes.data, ok.ctypes.data))
        return pred, best, ok, mu

    # ------------------------------------------------------------------ N1: Gaussian-MAP detector
    def cov_accumulate(self, Z: torch.Tensor, label: torch.Tensor, mean: torch.Tensor, k_sel: int = -1,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """float64 second moments of centred latents (08b:60-81, :276-296); ``k_sel < 0`` pools all classes (LDA)."""
        Z = self._dev(Z, torch.float32, "Z")
        label = self._dev(label, torch.int32, "label")
        mean = self._dev(mean, torch.float32, "mean")
        K, D = mean.shape
        if out is None:
            out = torch.zeros(D, D, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.avld_cov_accumulate(self._h, _ptr(Z), _ptr(label), _ptr(mean), int(k_sel), _ptr(out),
                                                Z.shape[0], K, D, _stream()))
        return out

    def map_score(self, Z: torch.Tensor, fit: MapFit, want_scores: bool = False, tau: Optional[float] = "fit"):
        """-> ``(pred [n] int32 index into fit.species or -1, best [n] f64, scores [n,K] f64 | None)``
        (09n:114-140, 10b:146-169)."""
        Z = self._dev(Z, torch.float32, "Z")
        K, D = fit.means.shape
        a, lp = fit.constants()
        dev = self.device
        mean = torch.from_numpy(np.ascontiguousarray(fit.means, np.float32)).to(dev)
        prec = torch.from_numpy(np.ascontiguousarray(fit.precision, np.float32)).to(dev)
        a_d, lp_d = torch.from_numpy(a).to(dev), torch.from_numpy(lp).to(dev)
        n = Z.shape[0]
        pred = torch.empty(n, dtype=torch.int32, device=dev)
        best = torch.empty(n, dtype=torch.float64, device=dev)
        scores = torch.empty(n, K, dtype=torch.float64, device=dev) if want_scores else None
        t = fit.tau if tau == "fit" else tau
        _lib.check(self.lib.avld_map_score(self._h, _ptr(Z), _ptr(mean), _ptr(prec), _ptr(a_d), _ptr(lp_d),
                                           float(t) if t is not None else 0.0, int(t is not None), _ptr(pred), _ptr(best),
                                           _ptr(scores), n, K, D, _stream()))
        return pred, best, scores

    def fit_map(self, Z: torch.Tensor, label: torch.Tensor, species_names: Sequence[str], **kw) -> MapFit:
        """See :func:`map_fit.fit_map` (08b_fit_map_detector.py:255-319)."""
        Z = self._dev(Z, torch.float32, "Z")
        label = self._dev(label, torch.int32, "label")
        return fit_map(self, Z, label, species_names, **kw)

    # ------------------------------------------------------------------ bring-up entry of the tcgen05 GEMM core
    def dbg_gemm(self, A: torch.Tensor, B: torch.Tensor, mode: int = 1) -> torch.Tensor:
        """``C[M,N] = A[M,K] @ B[N,K]^T`` through the split-precision tcgen05 core (tests only)."""
        A = self._dev(A, torch.float32, "A")
        B = self._dev(B, torch.float32, "B")
        M, K = A.shape
        N = B.shape[0]
        Cm = torch.empty(M, N, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.avld_dbg_gemm(self._h, _ptr(A), _ptr(B), _ptr(Cm), M, N, K, mode, _stream()))
        return Cm

    # ------------------------------------------------------------------ the fit (08:530-558), grid-aware, multi-GPU
    def fit_radial(self, Z: torch.Tensor, label: torch.Tensor, K: int, q_in: float = 0.95,
                   q_out: float | Sequence[float] = 0.01, *, group=None, semantics: str = "numpy2",
                   shard_rows: Optional[int] = None) -> RadialFit:
        """See :func:`radial_fit.fit_radial` (this engine supplies the CUDA kernels)."""
        Z = self._dev(Z, torch.float32, "Z")
        label = self._dev(label, torch.int32, "label")
        return fit_radial(self, Z, label, K, q_in, q_out, group=group, semantics=semantics, shard_rows=shard_rows)


def priority_ranks(species: Sequence[str], priority: Sequence[str]) -> np.ndarray:
    """Rank of each species under 09:428-436: first the species listed in ``PRIORITY_ORDER`` (in that
    order), then every other name in ``sorted()`` order."""
    rest = sorted(s for s in species if s not in priority)
    order = [s for s in priority if s in species] + rest
    return np.array([order.index(s) for s in species], dtype=np.int32)
