#!/usr/bin/env bash
# DFT GEMM bring-up probes: AVLD_DBG upper bounds (1 = no epilogue math, 2 = A operand always L2 resident, 3 = both).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for d in ${DBG_LIST:-0 1 2 3}; do
  AVLD_DBG=$d timeout 200 python bench.py --chunks 4096 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_dbg$d.log 2>&1
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_dbg$d.log").read().strip().splitlines()[-1])
print("dbg$d dft avg_launch_ms", d["roofline"]["avg_launch_ms"], "value", d["value"], d["clocks"])
PY
done
