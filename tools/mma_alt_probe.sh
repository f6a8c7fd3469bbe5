#!/usr/bin/env bash
cd "$(dirname "$0")/.."
for d in 5 13; do
  AVLD_DBG=$d timeout 200 python bench.py --chunks 4096 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_alt$d.log 2>&1
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_alt$d.log").read().strip().splitlines()[-1])
ms = d["roofline"]["avg_launch_ms"]
print("dbg=$d dft ms", ms, "cycles per MMA at 1965 MHz", ms * 1e-3 * 1.965e9 / (21 * 64 * 12))
PY
done
