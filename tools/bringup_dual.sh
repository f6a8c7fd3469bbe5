#!/usr/bin/env bash
# Bring-up of dftf4.cu (AVLD_DFT_DUAL=1: dual tile; =2: dual tile + B multicast in 4-CTA clusters) on a B200: parity first (bit-identical features expected), then an A/B of the
# STFT GEMM time on the same 8192-chunk workload.  Each step under its own timeout: a barrier-protocol bug shows up as a
# hang, and mbar_wait traps after its spin limit instead of spinning forever.
# Usage (one gpurun call):  gpurun --timeout 900 -- tools/bringup_dual.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# a failing mode can be excluded with AVLD_TEST_DUAL_MODES=1 (or =2)
AVLD_TEST_DUAL=1 timeout 300 python -m pytest tests/test_gpu_features.py -m gpu -q -x -k dual_tile > gpurun_out/dual_test.log 2>&1
echo "dual test rc $?"; tail -n 15 gpurun_out/dual_test.log
small="--chunks 8192 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 python bench.py $small > gpurun_out/dual_bench_base.log 2>&1; echo "base rc $?"
AVLD_DFT_DUAL=1 timeout 300 python bench.py $small > gpurun_out/dual_bench_dual.log 2>&1; echo "dual rc $?"
AVLD_DFT_DUAL=2 timeout 300 python bench.py $small > gpurun_out/dual_bench_mc.log 2>&1; echo "dual + multicast rc $?"
python - <<'PY'
import json
for tag in ("base", "dual", "mc"):
    try:
        line = [l for l in open(f"gpurun_out/dual_bench_{tag}.log") if l.startswith("{")][-1]
        d = json.loads(line)
        print(tag, "chunks/s", round(d["value"]), "dft ms/launch", round(d["roofline"]["avg_launch_ms"], 4),
              "stage ms", d.get("stage_ms_per_step", {}).get("gemm3_kernel<DFT>"))
    except Exception as e:
        print(tag, "no bench line:", e)
PY
