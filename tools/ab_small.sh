#!/usr/bin/env bash
# quick A/B of a kernel change: the small bench (16k chunks, stage times) with the product library and, when given, with
# environment switches of the bring-up build.  Usage: tools/ab_small.sh [VAR=1 ...]
cd "$(dirname "$0")/.."
small="--chunks 16384 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
show='import json,sys
d=json.loads(sys.stdin.read()); print(round(d["value"]), d["clocks"]["sm_mhz"], {k: round(v/16, 4) for k, v in d["stage_ms_per_step"].items() if v > 0.5})'
echo -n "product: "; python bench.py $small 2>&1 | tail -1 | python -c "$show"
for kv in "$@"; do
  echo -n "bringup $kv: "; env AVLD_LIB_PATH=amphibian_vae_latent_detector_b200/libavld_bringup.so $kv python bench.py $small 2>&1 | tail -1 | python -c "$show"
done
