#!/usr/bin/env python
"""Dense 16-bit tensor peaks of this B200 the way MEASURED_PEAKS.json is made (torch.matmul 8192^3: best of 10 = burst,
back to back for 4 s = sustained), for fp16 AND bf16: the STFT kernel multiplies fp16 operands, the encoder bf16.
Writes profiles/r02_dense_peaks.json (run on the GPU box, copy the file from gpurun_out/)."""
import json
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
N = 8192
out = {"how": "torch.matmul N=8192 (2 N^3 flop): best of 10 (burst), back to back for 4 s (sustained); CUDA events", "gpu": torch.cuda.get_device_name(0)}
for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
    a = torch.randn(N, N, device="cuda", dtype=dt)
    b = torch.randn(N, N, device="cuda", dtype=dt)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * N ** 3 / (e0.elapsed_time(e1) / 1e3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps, t0 = 0, time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < 4.0:
        for _ in range(20):
            a @ b
        reps += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    out[f"{name}_tflops"] = round(best, 1)
    out[f"{name}_tflops_sustained"] = round(2.0 * N ** 3 * reps / (e0.elapsed_time(e1) / 1e3) / 1e12, 1)
(REPO / "gpurun_out").mkdir(exist_ok=True)
(REPO / "gpurun_out" / "r02_dense_peaks.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out))
