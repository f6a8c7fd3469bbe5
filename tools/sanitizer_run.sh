#!/usr/bin/env bash
# compute-sanitizer over tools/sanitizer_workload.py: memcheck, then racecheck and synccheck (shared-memory hazards and
# barrier misuse in the hand-written kernels).  Summaries land in gpurun_out/sanitizer_<tool>.txt.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_workload.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "$tool rc $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer workload done|Error:" gpurun_out/sanitizer_$tool.txt | head -5
done
