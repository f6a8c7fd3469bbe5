#!/usr/bin/env python
"""Headline benchmark: end-to-end encode+detect chunks/s (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU algorithm on the host cores

Workload (BASELINE.json configs[1], weak-scaled for N > 1): per GPU `--chunks` (default 100 000) synthetic mono
3 s chunks (48 kHz, 144 000 samples), random-init stand-in encoder (seed 123).
The configs[2] gate -- the sharded fit's centroids / thresholds against a single-rank refit of the gathered latents (1e-6)
and bit-identity across ranks -- is evaluated after the timed region in every multi-rank run (`verify_fit` in the line);
`--verify-fit` also makes a failure fatal (run it at `--chunks 125000 --gpus 8` for the 1 M-chunk case).
One *step* = one pass of the whole hot path over the rank's resident chunks:
  RMS normalise (+PCM_16 round trip) -> STFT/mel/log/z-score -> encoder mu        [avld_encode]
  -> per-species centroid sums (+ all-reduce) -> radii (+ all-gather) -> exact q_in / q_out-grid quantiles
  -> accept / priority decision -> per-class decision histogram read back         [fit_radial + avld_decide]
`value` times that with inputs resident in HBM; `e2e` times the host-buffer C-ABI call
(avld_encode_detect_host_pcm16: pinned host audio in, decisions out, H2D/D2H inside the timed region); with N > 1 the host
slabs of a step form one pool that the ranks drain dynamically (a GPU behind a slower PCIe root takes fewer slabs), and a
concurrent raw-copy probe states the box's H2D ceiling next to it.
Inputs (57.6 GB per GPU) are far larger than the 126 MB L2, so no explicit L2 flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

SPECIES = ["Batrachyla_leptopus", "Batrachyla_taeniata", "Calyptocephalella_gayi", "Pleurodema_thaul"]
PRIORITY_ORDER = SPECIES                      # 09_evaluate_wav_detection.py:61-66
MEL_KW = dict(sr=48000, n_mels=64, fmin=150.0, fmax=15000.0, hop_length=384, n_fft=2048, target_frames=192)
CHUNK_LEN = 144000
Q_IN, Q_OUT_GRID = 0.95, (0.10, 0.15, 0.20, 0.25)          # run_qout_grid.sh:6, :13
DFT_FLOP_PER_CHUNK = 2.0 * 376 * 2048 * (2 * 634)          # 1.953 GFLOP (algorithmic, 1 pass)
METRIC = "encode+detect chunks/sec"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chunks", type=int, default=100000, help="resident chunks per GPU (one step processes all)")
    ap.add_argument("--e2e-chunks", type=int, default=16384, help="host-resident chunks per e2e step")
    ap.add_argument("--max-batch", type=int, default=1024, help="chunks per internal kernel pass")
    ap.add_argument("--cpu-chunks", type=int, default=1000, help="chunks of the CPU baseline sample (also the parity subsample)")
    ap.add_argument("--verify-fit", action="store_true", help="configs[2] gate: sharded fit == single-rank fit of the gathered latents")
    ap.add_argument("--scalar-semantics", default="numpy2", choices=["numpy1", "numpy2"],
                    help="rounding of rms_normalize's scalar arithmetic: numpy >= 2 (float32) or the reference's pinned numpy 1.26 (float64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_peaks():
    """HBM copy bandwidth and dense 16-bit tensor throughput of this pool's B200s.  MEASURED_PEAKS.json is written by the
    driver (bf16); profiles/r02_dense_peaks.json (tools/measure_dense_peaks.py, same method) adds the fp16 figure, which
    is the one the fp16 STFT kernel is held against."""
    out = dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        out.update(hbm_gbs=float(d["hbm_gbs"]), bf16_burst=float(d["bf16_tflops"]),
                   bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="MEASURED_PEAKS.json")
    out["fp16_sustained"], out["fp16_source"] = out["bf16_sustained"], out["source"] + " (bf16 figure: no fp16 measurement found)"
    q = REPO / "profiles" / "r02_dense_peaks.json"
    if q.exists():
        try:
            d = json.loads(q.read_text())
            out["fp16_sustained"], out["fp16_source"] = float(d["fp16_tflops_sustained"]), "profiles/r02_dense_peaks.json (torch.matmul fp16 8192^3, 4 s back to back)"
        except Exception:
            pass
    out["fp32_fma_tflops"] = 148 * 128 * 2 * 1.965e9 / 1e12          # 148 SMs x 128 lanes x 2 flop x 1.965 GHz ...
    try:                                                             # ... which a pure FFMA2 stream reaches (tools/fma_peak.cu)
        out["fp32_fma_tflops"] = float(json.loads((REPO / "profiles" / "r02j_fma_peak.json").read_text())["ffma2_tflops"])
    except Exception:
        pass
    return out


def traffic_of(chunks_per_launch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/dominant_kernel_traffic.json); null when the capture is of another kernel / size."""
    p = REPO / "profiles" / "dominant_kernel_traffic.json"
    try:
        d = json.loads(p.read_text())
    except Exception:
        return None, None
    if d.get("kernel") != "dftf3_kernel" or int(d.get("chunks_per_launch", 0)) != int(chunks_per_launch):
        return None, None
    return d["traffic_bytes"], d.get("source")


def place_device(local_rank: int, world: int):
    """Which visible GPU a rank drives.  With more GPUs visible than ranks, the ranks are spread over the two halves of the
    PCI bus order: on this pool's 8-GPU VM four GPUs share one host bridge whose pinned-copy ceiling is ~115 GB/s in total
    (profiles/r02_h2d_multirank_probe_8gpu.json: 2 GPUs of a half copy at 55.6 GB/s each, 4 at 28.8), so a 4-rank job that
    takes two GPUs from each half keeps the full PCIe rate per GPU.  Identity when every visible GPU has a rank."""
    import torch
    ndev = torch.cuda.device_count()
    if world <= 1 or ndev <= world:
        return local_rank, "identity"
    order = sorted(range(ndev), key=lambda i: (int(getattr(torch.cuda.get_device_properties(i), "pci_bus_id", i)), i))
    lower, upper = order[: ndev // 2], order[ndev // 2:]
    spread = [g for pair in zip(upper, lower) for g in pair] + upper[len(lower):]
    return spread[local_rank], f"spread over PCI halves {spread[:world]}"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------
# CPU legs (oracle = numpy/torch restatement of the reference algorithm; oracle/ is test infrastructure and is
# only *timed* here, never used by the CUDA path)
# ------------------------------------------------------------------------------------------------------------
def _cpu_fit_detect(Z, labels, hp):
    cent, rk, _, _ = hp.fit_radial(Z, labels, 4, Q_IN, Q_OUT_GRID[0])
    for q in Q_OUT_GRID[1:]:
        hp.fit_radial(Z, labels, 4, Q_IN, q)
    pred, best, _ = hp.decide_batch(Z, SPECIES, cent, rk)
    return pred


def cpu_reference_loop(x: np.ndarray, labels: np.ndarray, numpy1: bool = False):
    """The reference's execution model: one chunk at a time, batch 1 (08:488-506, 10:395-418)."""
    from oracle import hotpath as hp
    from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder
    enc = build_standin_encoder(seed=123)
    t0 = time.perf_counter()
    y, ok, _ = hp.rms_normalize_batch(x, pcm16=True, numpy1_scalars=numpy1)
    t1 = time.perf_counter()
    feats = [hp.logmel_features(row, **MEL_KW) for row in y]                  # M1-M5, one chunk at a time
    t2 = time.perf_counter()
    Z = np.stack([hp.encode_features(enc, f) for f in feats])                 # E0-E2, batch 1
    t3 = time.perf_counter()
    _cpu_fit_detect(Z, labels, hp)
    t4 = time.perf_counter()
    n = x.shape[0]
    stages = {"normalise": n / (t1 - t0), "features": n / (t2 - t1), "encoder": n / (t3 - t2), "fit+detect": n / (t4 - t3)}
    return t4 - t0, {k: round(v, 1) for k, v in stages.items()}, (Z, ok)


def parity_on_sample(eng, x_dev, labels_np, Zo, oko):
    """The CPU leg's outputs double as the checker (parity on a 1 000-chunk subsample of the benchmark's own
    workload): latents of the CUDA path against the oracle's, thresholds of a fit on either, decisions outside a 1e-3 band."""
    import torch
    from oracle import hotpath as hp
    from amphibian_vae_latent_detector_b200.engine import priority_ranks
    Z, ok = eng.encode(x_dev, pcm16=True)
    Zg = Z.cpu().numpy()
    lat_err = float(np.max(np.abs(Zg - Zo)) / np.max(np.abs(Zo)))
    cent_o, rk_o, _, _ = hp.fit_radial(Zo, labels_np, 4, Q_IN, Q_OUT_GRID[-1])
    fit = eng.fit_radial(Z, torch.from_numpy(labels_np).to(Z.device), 4, Q_IN, (Q_OUT_GRID[-1],))
    thr_err = float(np.nanmax(np.abs(fit.rk[0] - rk_o) / np.abs(rk_o)))
    pred_o, _, radii_o = hp.decide_batch(Zo, SPECIES, cent_o, rk_o)
    prio = torch.from_numpy(priority_ranks(SPECIES, PRIORITY_ORDER)).to(Z.device)
    pred, _ = eng.decide(fit.radii_local, torch.from_numpy(fit.rk[0]).to(Z.device), prio)
    pred = pred.cpu().numpy()
    near = np.any(np.abs(radii_o - rk_o[None]) / rk_o[None] <= 1e-3, axis=1)
    res = {"chunks": int(Zo.shape[0]), "gate_equal": bool(np.array_equal(ok.cpu().numpy(), oko)), "latent_max_rel_err": lat_err,
           "threshold_max_rel_err": thr_err, "decisions_equal_outside_1e-3_band": bool(np.array_equal(pred[~near], pred_o[~near])),
           "chunks_inside_band": int(near.sum()), "tolerance": 1e-3}
    res["ok"] = bool(res["gate_equal"] and lat_err <= 1e-3 and thr_err <= 1e-3 and res["decisions_equal_outside_1e-3_band"])
    return res


_W = {}


def _pool_init(per_worker):
    """Each worker synthesises its own chunks once (setup, untimed) and builds the encoder."""
    import torch
    torch.set_num_threads(1)
    try:                                   # numpy's BLAS must not spawn a thread team per worker either
        import threadpoolctl
        _W["blas_limit"] = threadpoolctl.threadpool_limits(1)
    except ImportError:
        pass
    from amphibian_vae_latent_detector_b200 import synth
    from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder
    x, _ = synth.make_chunks(per_worker, CHUNK_LEN, seed=123, first_index=(os.getpid() % 9973) * per_worker)
    _W["x"] = x.numpy()
    _W["enc"] = build_standin_encoder(seed=123)


def _pool_worker(_):
    from oracle import hotpath as hp
    y, ok, _r = hp.rms_normalize_batch(_W["x"], pcm16=True)
    return hp.encode_batch(_W["enc"], y, **MEL_KW)


def workload_config(chunks_per_gpu, world, **extra):
    cfg = {"workload": "100k synthetic 3 s mono chunks per GPU (configs[1]): RMS-normalise + log-mel + encoder mu + "
                       "radial fit (q_in 0.95, q_out grid) + decision; stand-in encoder, random init",
           "chunk_len": CHUNK_LEN, "chunks_per_gpu": chunks_per_gpu, "parallelism": f"dp{world}"}
    cfg.update(extra)
    return cfg


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU algorithm (numpy/torch oracle port; the reference itself is Python
    and cannot travel to the GPU box) on all host cores: a process pool, one chunk at a time inside each worker,
    each step a bounded sample of the workload.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import hotpath as hp
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    per_worker = 32
    n_step = workers * per_worker
    # One thread per worker, one worker per core.  Without these the BLAS behind numpy starts a full thread team in
    # every worker (cores x cores spinning threads) and the arm runs ~15x slower than the cores allow.
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    ctx = mp.get_context("spawn")
    times = []
    labels = (np.arange(n_step) % 4).astype(np.int32)
    with ctx.Pool(workers, initializer=_pool_init, initargs=(per_worker,)) as pool:
        pool.map(_pool_worker, range(workers), chunksize=1)          # spin-up
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_pool_worker, range(workers), chunksize=1)
            Z = np.concatenate(res)
            _cpu_fit_detect(Z, labels, hp)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = n_step * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 FFT)",
            "data": "synthetic",
            "config": workload_config(args.chunks, args.gpus, sample_chunks_per_step=n_step),
            "cpu_baseline": {"value": value, "unit": "chunks/s", "cores": workers, "kind": "port",
                             "sample": f"{n_step} chunks per step: process pool of {workers} single-thread workers x "
                                       f"{per_worker} chunks, numpy/torch oracle of the reference algorithm, batch 1"},
            "e2e": {"value": value, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------------------
def kernel_rooflines(stages, steps, chunks_timed, info, peaks, max_batch):
    """achieved / peak / frac for every kernel family that takes >= 2 % of the step.  Algorithmic work per chunk (3 s, 376
    frames, 634 mel-weighted bins): bytes a perfect implementation must move, or flops it must issue."""
    L, F, M, T = CHUNK_LEN, 376, 64, 192
    enc_conv_flops = 2.0 * 9 * (96 * 32 * 32 * 64 + 48 * 16 * 64 * 128 + 24 * 8 * 128 * 128) * 4      # convs 2-4 (pre-pool pixels)
    spec = {
        "prep_kernel": ("hbm", 4 * L + 2 * L, "reads the float32 chunk once (4 L), writes the normalised PCM_16 integers (2 L)"),
        "fold3_kernel": ("hbm", 2 * L + 4 * F * 2048, "reads the integers (2 L), writes the folded fp16 hi / lo tiles (4 B x 376 x 2048)"),
        "dftf3_kernel": ("tensor_fp16", info["issued_flops_per_chunk"], "issued split-precision flops (3 passes) of the folded DFT GEMM"),
        "logmel_post_kernel": ("hbm", 3 * F * M * 4 * 2 + T * M * 4, "reads and clears three mel-power planes, writes the [192, 64] features"),
        "conv1_kernel": ("fma_fp32", 2.0 * 9 * 32 * 192 * 64, "3x3 conv, 1 -> 32 channels on 192 x 64 pixels (CUDA cores)"),
        "convh_kernel": ("tensor_bf16", 3.0 * enc_conv_flops / 4, "3x3 convs 32->64, 64->128, 128->128, three bf16 passes"),
        "gemm3_kernel": ("tensor_bf16", 3.0 * 2.0 * (6144 * 512 + 512 * 128), "dense 6144->512->128, three bf16 passes"),
    }
    total_ms = sum(v["ms"] for v in stages.values())
    out = []
    for name, (bound, work, what) in spec.items():
        st = stages.get(name)
        if not st or st["ms"] <= 0 or st["ms"] < 0.02 * total_ms:
            continue
        sec = st["ms"] / 1e3
        if bound == "hbm":
            ach, peak, unit = work * chunks_timed / sec / 1e9, peaks["hbm_gbs"], "GB/s"
        else:
            ach, unit = work * chunks_timed / sec / 1e12, "TFLOP/s"
            peak = {"tensor_fp16": peaks["fp16_sustained"], "tensor_bf16": peaks["bf16_sustained"], "fma_fp32": peaks["fp32_fma_tflops"]}[bound]
        out.append({"kernel": name, "bound": bound, "achieved": round(ach, 1), "peak": round(peak, 1), "unit": unit,
                    "frac": round(ach / peak, 3), "ms_per_launch": round(st["ms"] / max(st["timed_launches"], 1), 4),
                    "share_of_step": round(st["ms"] / total_ms, 3), "work_per_chunk": work, "what": what})
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from amphibian_vae_latent_detector_b200 import synth
    from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder
    from amphibian_vae_latent_detector_b200.engine import Engine, priority_ranks

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CUDA path has no CPU fallback")
    gpu, placement = place_device(local_rank, world)
    torch.cuda.set_device(gpu)
    dev = torch.device("cuda", gpu)
    group, store = None, None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
        store = dist.distributed_c10d._get_default_store()

    def barrier():
        if world > 1:
            dist.barrier(group=group, device_ids=[gpu])
        torch.cuda.synchronize()

    # ---------------- setup (untimed)
    n = args.chunks
    eng = Engine(gpu, chunk_len=CHUNK_LEN, max_batch=args.max_batch, scalar_semantics=args.scalar_semantics)
    eng.load_encoder(build_standin_encoder(seed=123))
    X = torch.empty(n, CHUNK_LEN, dtype=torch.float32, device=dev)
    labels = torch.empty(n, dtype=torch.int32, device=dev)
    slab = 1024
    for i in range(0, n, slab):
        m = min(slab, n - i)
        xs, ls = synth.make_chunks(m, CHUNK_LEN, seed=123, first_index=rank * n + i, device=dev)
        X[i:i + m], labels[i:i + m] = xs, ls
    del xs, ls
    torch.cuda.empty_cache()
    prio = priority_ranks(SPECIES, PRIORITY_ORDER)
    prio_d = torch.from_numpy(prio).to(dev)
    state = {}

    def step():
        Z, ok = eng.encode(X, pcm16=True)
        # every rank holds n rows: the gather block size is known, the fit's device work runs without a host sync
        fit = eng.fit_radial(Z, labels, 4, Q_IN, Q_OUT_GRID, group=group, shard_rows=n)
        thr = torch.from_numpy(fit.rk[0]).to(dev)
        pred, best = eng.decide(fit.radii_local, thr, prio_d)
        hist = torch.bincount((pred + 1).long(), minlength=5).cpu()        # D2H read of the step's result
        state.update(fit=fit, hist=hist, ok=ok, Z=Z)
        return hist

    for _ in range(args.warmup):
        step()
    eng.collect(reset=True)
    eng.profile(True)
    sampler = ClockSampler(gpu)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    eng.profile(False)
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=group)
    total_ms = float(ms.item())
    value = world * n * args.steps / (total_ms / 1e3)
    stages = eng.collect(reset=True)
    launches = sum(v["launches"] for v in stages.values())
    info = eng.dft_info()

    # ---------------- configs[2] gate: the sharded fit against a single-rank refit of all latents
    verify = None
    if args.verify_fit or world > 1:          # multi-rank runs always carry the gate's outcome in their line
        fit, Z = state["fit"], state["Z"]
        if world > 1:
            Zs = [torch.empty_like(Z) for _ in range(world)] if rank == 0 else None
            Ls = [torch.empty_like(labels) for _ in range(world)] if rank == 0 else None
            dist.gather(Z, Zs, dst=0, group=group)
            dist.gather(labels, Ls, dst=0, group=group)
            mine = torch.from_numpy(np.concatenate([fit.rk.ravel(), fit.rk_in, fit.centroids.ravel().astype(np.float64)])).to(dev)
            alls = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(alls, mine, group=group)
            identical = all(torch.equal(torch.nan_to_num(a), torch.nan_to_num(alls[0])) for a in alls)
        else:
            Zs, Ls, identical = [Z], [labels], True
        if rank == 0:
            ref = eng.fit_radial(torch.cat(Zs), torch.cat(Ls), 4, Q_IN, Q_OUT_GRID)
            thr_err = float(np.nanmax(np.abs(fit.rk - ref.rk) / np.abs(ref.rk)))
            cen_err = float(np.nanmax(np.abs(fit.centroids - ref.centroids)) / np.nanmax(np.abs(ref.centroids)))
            verify = {"total_chunks": int(world * n), "ranks": world, "threshold_max_rel_err_vs_single_rank": thr_err,
                      "centroid_max_rel_err_vs_single_rank": cen_err, "counts_equal": bool(np.array_equal(fit.counts, ref.counts)),
                      "bit_identical_across_ranks": bool(identical), "tolerance": 1e-6}
            verify["ok"] = bool(thr_err <= 1e-6 and cen_err <= 1e-6 and verify["counts_equal"] and identical)
            del ref
        del Zs, Ls
        barrier()
    state.pop("Z", None)

    # ---------------- e2e through the host-buffer C-ABI calls (pinned host audio in, decisions out)
    e2e = None
    e2e_f32 = None
    if not args.no_e2e:
        ne = min(args.e2e_chunks, n)
        fit = state["fit"]
        cent, thr = np.nan_to_num(fit.centroids), fit.rk[0]
        # chunks per host call = one unit of the dynamic pool.  One or two ranks copy at the full PCIe rate each and have nothing to
        # balance: a rank's step is one call (the un-overlapped tail of a call, ~1 ms of kernels after the last byte has landed, is
        # paid once per call); more ranks share the host bridges and balance in 4096-chunk units
        grab = min(ne if world <= 2 else 4096, ne)
        run_id = [0]

        def run_e2e(xh, bytes_per_sample, api):
            """`steps` x world x rows chunks of host audio as ONE pool of `grab`-chunk slabs; a rank takes the next slab when
            it has finished its last (atomic counter in the rendezvous store), so the split follows what each GPU's PCIe
            path delivers.  Time = barrier to barrier, max over ranks."""
            rows = xh.shape[0]
            per_rank = rows // grab
            n_units = per_rank * world * args.steps
            for _ in range(2):
                eng.encode_detect_host(xh[:grab], cent, thr, prio, pcm16=True)
            run_id[0] += 1
            key = f"e2e_pool_{run_id[0]}"
            barrier()
            t0 = time.perf_counter()
            done = 0
            while True:
                u = (store.add(key, 1) - 1) if store is not None else done
                if u >= n_units:
                    break
                j = u % per_rank
                eng.encode_detect_host(xh[j * grab:(j + 1) * grab], cent, thr, prio, pcm16=True)
                done += 1
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            mine = torch.tensor([float(done)], dtype=torch.float64, device=dev)
            shares = [mine]
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX, group=group)
                shares = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(shares, mine, group=group)
            sec = float(dt.item())
            total_chunks = n_units * grab
            h2d = total_chunks * CHUNK_LEN * bytes_per_sample // (args.steps * world)
            return {"value": total_chunks / sec, "unit": "chunks/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(total_chunks // (args.steps * world)) * (4 + 4 + 1),
                    "chunks_per_step": int(total_chunks // (args.steps * world)), "chunks_per_host_call": int(grab),
                    "h2d_gbs_per_gpu": total_chunks * CHUNK_LEN * bytes_per_sample / sec / 1e9 / world,
                    "slabs_taken_per_rank": [int(s.item()) for s in shares], "api": api}

        def h2d_ceiling(xh):
            """All ranks copy their pinned slab to the device at once, no kernels: the box's concurrent H2D rate per GPU."""
            dst = torch.empty_like(xh[:grab], device=dev)
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                dst.copy_(xh[:grab], non_blocking=True)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 12
            with torch.cuda.stream(s):
                a.record()
                for r in range(reps):
                    dst.copy_(xh[(r % (xh.shape[0] // grab)) * grab:][:grab], non_blocking=True)
                b.record()
            torch.cuda.synchronize()
            gbs = torch.tensor([dst.numel() * dst.element_size() * reps / (a.elapsed_time(b) / 1e3) / 1e9], dtype=torch.float64, device=dev)
            alls = [gbs]
            if world > 1:
                alls = [torch.zeros_like(gbs) for _ in range(world)]
                dist.all_gather(alls, gbs, group=group)
            barrier()
            return [round(float(v.item()), 2) for v in alls]

        # (a) PCM_16 samples, the format of the reference's WAV datasets (00:57 writes PCM_16, core:210 reads it)
        xh16 = torch.empty(ne, CHUNK_LEN, dtype=torch.int16, pin_memory=True)
        for i in range(0, ne, 2048):
            m = min(2048, ne - i)
            xh16[i:i + m].copy_(torch.clamp(torch.round(X[i:i + m] * 32767.0), -32768, 32767).to(torch.int16))
        torch.cuda.synchronize()
        e2e = run_e2e(xh16, 2, "avld_encode_detect_host_pcm16 (pinned host PCM_16 audio -> decisions)")
        ceil = h2d_ceiling(xh16)
        e2e["h2d_ceiling_gbs_per_gpu"] = ceil
        e2e["h2d_ceiling_gbs_total"] = round(sum(ceil), 1)
        e2e["frac_of_ceiling"] = round(e2e["h2d_gbs_per_gpu"] * world / sum(ceil), 3)
        del xh16
        # (b) float32 samples (twice the PCIe bytes)
        nf = min(ne, 8192)
        xh = torch.empty(nf, CHUNK_LEN, dtype=torch.float32, pin_memory=True)
        xh.copy_(X[:nf])
        torch.cuda.synchronize()
        e2e_f32 = run_e2e(xh, 4, "avld_encode_detect_host (pinned host float32 audio -> decisions)")
        del xh
        eng.collect(reset=True)

    if rank == 0:
        peaks = measured_peaks()
        dft = stages["dftf3_kernel"]
        chunks_timed = n * args.steps
        dft_tflops = DFT_FLOP_PER_CHUNK * chunks_timed / (dft["ms"] / 1e3) / 1e12 if dft["ms"] > 0 else None
        stage_ms = {k: round(v["ms"] / args.steps, 3) for k, v in stages.items() if v["timed_launches"]}
        kernel_ms = sum(stage_ms.values())
        issued_tflops = (info["issued_flops_per_chunk"] * chunks_timed / (dft["ms"] / 1e3) / 1e12) if dft["ms"] > 0 else None
        traffic, traffic_src = traffic_of(args.max_batch)
        roofline = {
            "bound": "tensor", "kernel": "dftf3_kernel (windowed DFT as GEMM, folded three times, cta_group::2, + |X|^2 + mel)",
            # algorithmic = the plain DFT GEMM over the 634 bins with mel weight (1.953 GFLOP per
            # chunk) / the kernel's measured launch time.  The folded kernel reaches the same bins with 0.57 x those tensor
            # flops, so `achieved` can exceed what the pipe executes: `issued_tflops` / `tensor_pipe_frac` is the
            # utilisation of the tensor pipe itself.
            "achieved": dft_tflops, "peak": peaks["fp16_sustained"], "unit": "TFLOP/s",
            "frac": (dft_tflops / peaks["fp16_sustained"]) if dft_tflops else None,
            "peak_source": f"dense fp16 sustained: {peaks['fp16_source']}",
            "algorithmic_gflop_per_chunk": DFT_FLOP_PER_CHUNK / 1e9,
            "issued_gflop_per_chunk": info["issued_flops_per_chunk"] / 1e9,
            "issued_over_algorithmic": info["issued_flops_per_chunk"] / DFT_FLOP_PER_CHUNK,
            "issued_tflops": issued_tflops,
            "tensor_pipe_frac": (issued_tflops / peaks["fp16_sustained"]) if issued_tflops else None,
            "note": "achieved / frac count the flops of the plain DFT GEMM over the 634 mel-weighted bins (SURVEY 8d); the folded "
                    "kernel reaches the same bins with issued_over_algorithmic x those flops, so frac can exceed 1 -- "
                    "tensor_pipe_frac (issued flops / peak) is the utilisation of the pipe",
            "dft_mode": info["mode"], "chunks_per_launch": args.max_batch,
            "avg_launch_ms": dft["ms"] / max(dft["timed_launches"], 1), "traffic": traffic, "traffic_source": traffic_src,
            "share_of_kernel_time": (stage_ms.get("dftf3_kernel", 0.0) / kernel_ms) if kernel_ms else None}
        line = {
            "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16x2 / bf16x2 split operands, fp32 accumulate (TMEM)", "data": "synthetic",
            "config": workload_config(n, world, max_batch=args.max_batch, device_placement=placement,
                                      rms_scalar_semantics=args.scalar_semantics,
                                      l2="inputs (57.6 GB/GPU at 100k chunks) exceed the 126 MB L2; no flush needed"),
            "clocks": clocks,
            "gpu_launches": int(launches),
            "roofline": roofline,
            "rooflines": kernel_rooflines(stages, args.steps, chunks_timed, info, peaks, args.max_batch),
            "stage_ms_per_step": stage_ms,
            "decision_hist": state["hist"].tolist(),
        }
        if verify is not None:
            line["verify_fit"] = verify
        if e2e is not None:
            line["e2e"] = e2e
            line["e2e_f32_host"] = e2e_f32
        if world == 1 and not args.no_cpu_baseline:
            nc = min(args.cpu_chunks, n)
            xc, lc = synth.make_chunks(nc, CHUNK_LEN, seed=123, first_index=0)
            dt, cpu_stages, (Zo, oko) = cpu_reference_loop(xc.numpy(), lc.numpy(), numpy1=args.scalar_semantics == "numpy1")
            line["cpu_baseline"] = {"value": nc / dt, "unit": "chunks/s", "cores": int(torch.get_num_threads()),
                                    "kind": "port", "stages_chunks_per_s": cpu_stages,
                                    "sample": f"{nc} chunks of the same workload, one chunk at a time (batch 1) as the "
                                              f"reference runs it: numpy float64 FFT + torch CPU encoder, "
                                              f"{dt:.1f} s wall, host has {os.cpu_count()} logical cores; one process, as the "
                                              f"reference executes -- all cores at once (process pool): --impl reference"}
            line["parity"] = parity_on_sample(eng, xc.to(dev), lc.numpy(), Zo, oko)  # the CUDA path on the CPU leg's own chunks
        print(json.dumps(line), flush=True)
        if args.verify_fit and verify is not None and not verify["ok"]:
            raise SystemExit(f"--verify-fit failed: {verify}")
        if "parity" in line and not line["parity"]["ok"]:
            raise SystemExit(f"parity check on the CPU sample failed: {line['parity']}")
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
