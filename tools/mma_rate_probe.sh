#!/usr/bin/env bash
# tcgen05.mma.cta_group::2 issue rate of the fold2 kernel with no operand loads and no epilogue math, against N
cd "$(dirname "$0")/.."
for n in 64 128 160 192 256; do
  AVLD_DBG=5 AVLD_DBG_N=$n timeout 200 python bench.py --chunks 4096 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_n$n.log 2>&1
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n$n.log").read().strip().splitlines()[-1])
ms = d["roofline"]["avg_launch_ms"]
mmas = 21 * 64 * 12          # per cluster: waves x stages x MMAs
print("N=$n dft ms", ms, "cycles per MMA at 1965 MHz", ms * 1e-3 * 1.965e9 / mmas)
PY
done
