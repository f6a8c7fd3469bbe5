#!/usr/bin/env bash
# Runs every GPU test file in its own process (a trapped kernel poisons the CUDA context of its
# process only) under a timeout, logging to gpurun_out/.  Usage: tools/gpu_checks.sh [file ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
files=("$@")
if [ ${#files[@]} -eq 0 ]; then
  files=(tests/test_gpu_gemm.py tests/test_gpu_rms.py tests/test_gpu_radial.py tests/test_gpu_features.py tests/test_gpu_encoder.py tests/test_gpu_e2e.py tests/test_gpu_reference_api.py tests/test_gpu_map.py tests/test_gpu_pipeline.py tests/test_gpu_stream.py tests/test_gpu_grid_c4.py tests/test_gpu_edge_cases.py tests/test_gpu_zmap_cli.py)
fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in "${files[@]}"; do
  name=$(basename "$f" .py)
  echo "=== $f"
  timeout 420 python -m pytest "$f" -m gpu -q -x --timeout 300 > "gpurun_out/$name.log" 2>&1
  r=$?
  tail -n 25 "gpurun_out/$name.log"
  echo "=== $f exit $r"
  [ $r -ne 0 ] && rc=1
done
exit $rc
