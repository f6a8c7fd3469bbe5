#!/usr/bin/env python
"""The C ABI's own collectives (avld_comm_init / avld_allreduce_centroids / avld_allgather_radii, NCCL loaded with dlopen)
on N GPUs of one box, no torch.distributed: every rank fits its ragged shard of synthetic latents through
fit_radial(group="avld"), rank 0 refits everything on one GPU, all results must be bit-identical.

    python tools/comm_check.py [n_gpus] [rows]          -> one JSON line (gpurun_out/comm_check.json)"""
import json
import sys
from pathlib import Path

import numpy as np
import torch
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))


def make(rows, D=128, K=4, seed=123):
    g = np.random.default_rng(seed)
    lab = g.integers(0, K, rows).astype(np.int32)
    cent = g.normal(size=(K, D)).astype(np.float32) * 2
    return (cent[lab] + g.normal(size=(rows, D)).astype(np.float32)), lab


def worker(rank, world, rows, q):
    from amphibian_vae_latent_detector_b200.engine import Engine
    torch.cuda.set_device(rank)
    eng = Engine(rank, chunk_len=144000, max_batch=4)
    if rank == 0:
        uid = eng.comm_unique_id()
        for _ in range(world - 1):
            q["id"].put(uid)
    else:
        uid = q["id"].get()
    eng.comm_init(uid, rank, world)
    Z, lab = make(rows)
    cuts = np.linspace(0, rows, world + 1).astype(int)
    cuts[1:-1] += np.arange(1, world) * 7          # ragged shards
    sl = slice(cuts[rank], cuts[rank + 1])
    shard_rows = int(np.max(np.diff(cuts)))
    Zd, ld = torch.from_numpy(Z[sl]).cuda(), torch.from_numpy(lab[sl]).cuda()
    grid = (0.10, 0.15, 0.20, 0.25)
    fit = eng.fit_radial(Zd, ld, 4, 0.95, grid, group="avld", shard_rows=shard_rows)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(5):
        eng.fit_radial(Zd, ld, 4, 0.95, grid, group="avld", shard_rows=shard_rows)
    ev1.record()
    torch.cuda.synchronize()
    out = {"rank": rank, "rk": fit.rk.tolist(), "rk_in": fit.rk_in.tolist(), "centroids": fit.centroids.tobytes().hex(),
           "counts": fit.counts.tolist(), "ms_per_fit": ev0.elapsed_time(ev1) / 5}
    if rank == 0:
        ref = eng.fit_radial(torch.from_numpy(Z).cuda(), torch.from_numpy(lab).cuda(), 4, 0.95, grid)
        out["single"] = {"rk": ref.rk.tolist(), "rk_in": ref.rk_in.tolist(), "centroids": ref.centroids.tobytes().hex(),
                         "counts": ref.counts.tolist()}
    q["out"].put(out)
    q["done"].get()          # keep the communicator alive until every rank has reported
    eng.close()


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
    ctx = mp.get_context("spawn")
    q = {"id": ctx.Queue(), "out": ctx.Queue(), "done": ctx.Queue()}
    procs = [ctx.Process(target=worker, args=(r, world, rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted((q["out"].get(timeout=300) for _ in range(world)), key=lambda o: o["rank"])
    for _ in range(world):
        q["done"].put(1)
    for p in procs:
        p.join(timeout=60)
    single = outs[0]["single"]
    same = all(o[k] == single[k] for o in outs for k in ("rk", "rk_in", "centroids", "counts"))
    line = {"what": "sharded radial fit through the C ABI's own NCCL collectives (no torch.distributed)", "gpus": world, "rows": rows,
            "bit_identical_to_single_gpu_fit_on_every_rank": bool(same), "ms_per_fit": max(o["ms_per_fit"] for o in outs),
            "rk_q_out_0.10": outs[0]["rk"][0]}
    (REPO / "gpurun_out").mkdir(exist_ok=True)
    (REPO / "gpurun_out" / "comm_check.json").write_text(json.dumps(line) + "\n")
    print(json.dumps(line))
    if not same:
        raise SystemExit("sharded fit differs from the single-GPU fit")


if __name__ == "__main__":
    main()
