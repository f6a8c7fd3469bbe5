#!/usr/bin/env python
"""BASELINE.json configs[2] (sharded fit): under torchrun with W ranks, every rank builds the same N latents, fits them
alone (the 1-GPU answer) and then fits only its contiguous shard with the NCCL exchanges (one all-reduce of fp64
sums/counts, one all-gather of radii).  The sharded thresholds must equal the 1-GPU ones (1e-6 rel.;
here they are bit-identical because the order statistics are exact and the centroid sums are fp64).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/multigpu_fit_check.py [N]
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from amphibian_vae_latent_detector_b200.engine import Engine  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K, D, GRID = 4, 128, (0.10, 0.15, 0.20, 0.25)
g = torch.Generator(device=dev).manual_seed(123)
cents = 3.0 * torch.randn(K, D, generator=g, device=dev)
label = (torch.arange(N, device=dev) % K).to(torch.int32)
Z = cents[label.long()] + torch.randn(N, D, generator=g, device=dev)
eng = Engine(local, chunk_len=144000, max_batch=8)
one = eng.fit_radial(Z, label, K, 0.95, GRID)
lo, hi = rank * N // world + (7 if rank else 0), (rank + 1) * N // world + (7 if rank + 1 < world else 0)   # ragged shards
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng.fit_radial(Z[lo:hi], label[lo:hi], K, 0.95, GRID, group=dist.group.WORLD)
dist.barrier(device_ids=[local]); torch.cuda.synchronize()
ev0.record()
sh = eng.fit_radial(Z[lo:hi], label[lo:hi], K, 0.95, GRID, group=dist.group.WORLD)
ev1.record(); torch.cuda.synchronize()
ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
same = {k: bool(np.array_equal(getattr(one, k), getattr(sh, k))) for k in ("rk_in", "rk_out", "rk", "counts")}
same["centroids"] = bool(np.array_equal(one.centroids, sh.centroids))
flags = torch.tensor([int(all(same.values()))], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "N": N, "K": K, "D": D, "q_out_grid": GRID, "sharded_fit_ms_max_over_ranks": float(ms.item()),
                      "bit_identical_to_1gpu": same, "all_ranks_agree": bool(flags.item()),
                      "rk": sh.rk.tolist()}))
dist.destroy_process_group()
