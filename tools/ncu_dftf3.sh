#!/usr/bin/env bash
# One `ncu --set full` capture of the STFT GEMM (source view included), after a plain run of the same command exited 0.
# Usage: gpurun -- tools/ncu_dftf3.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-r02}
mkdir -p gpurun_out
one="--chunks 2048 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 python bench.py $one > gpurun_out/bench_one_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/bench_one_$tag.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dftf3 -c 1 -f -o gpurun_out/prof_dftf3_$tag \
  python bench.py $one > gpurun_out/ncu_dftf3_$tag.log 2>&1; echo "ncu dftf3 rc $?"
