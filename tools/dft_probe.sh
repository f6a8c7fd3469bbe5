#!/usr/bin/env bash
# DFT GEMM bring-up probes: pipeline trace of cluster 0 and the AVLD_DBG upper bounds
# (1 = no epilogue math, 2 = operands always L2/smem resident, 3 = both).  Logs into gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AVLD_TRACE=gpurun_out/trace_new.txt timeout 120 python tools/trace_run.py > gpurun_out/trace_run.log 2>&1
for d in 0 1 2 3; do
  AVLD_DBG=$d timeout 200 python bench.py --chunks 4096 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_dbg$d.log 2>&1
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_dbg$d.log").read().strip().splitlines()[-1])
print("dbg$d dft avg_launch_ms", d["roofline"]["avg_launch_ms"], "value", d["value"])
PY
done
