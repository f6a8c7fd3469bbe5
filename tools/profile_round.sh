#!/usr/bin/env bash
# Round profile: full GPU test suite, the default bench (plain), then the ncu launch list and --set full captures.
# Numbers printed under ncu are never bench values.  Usage: tools/profile_round.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-r01}
mkdir -p gpurun_out
tools/gpu_checks.sh > gpurun_out/checks_$tag.log 2>&1; echo "checks rc $?"
timeout 900 python bench.py > gpurun_out/bench_$tag.log 2>&1; echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.log 2>&1; echo "bench ref rc $?"
small="--chunks 8192 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 python bench.py $small > gpurun_out/bench_small_$tag.log 2>&1; echo "bench small rc $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'prep_|fold|dftf|logmel_post|conv|gemm3|radii_kernel|decide_kernel|centroid_kernel|select_hist|split_' -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py $small > gpurun_out/ncu_launches_$tag.log 2>&1; echo "ncu launches rc $?"
one="--chunks 2048 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dftf3 -c 1 -f -o gpurun_out/prof_dftf3_$tag \
  python bench.py $one > gpurun_out/ncu_dftf3_$tag.log 2>&1; echo "ncu dftf3 rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fold3|prep_|logmel_post|conv1|convh|gemm3' -c 10 -f \
  -o gpurun_out/prof_stream_$tag python bench.py $one > gpurun_out/ncu_stream_$tag.log 2>&1; echo "ncu stream rc $?"
tail -c 400 gpurun_out/bench_$tag.log
