#!/usr/bin/env bash
# A/B of two builds of the library on the small bench (stage times per 16 passes), alternating, three rounds:
#   tools/ab_lib.sh <other.so>      (the product libavld.so against another build selected through AVLD_LIB_PATH)
cd "$(dirname "$0")/.."
other=${1:-amphibian_vae_latent_detector_b200/libavld_prev.so}
small="--chunks 16384 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
show='import json,sys
d=json.loads(sys.stdin.read()); print(round(d["value"]), d["clocks"]["sm_mhz"], {k: round(v, 3) for k, v in d["stage_ms_per_step"].items() if v > 0.5})'
for i in 1 2 3; do
  echo -n "product: "; python bench.py $small 2>&1 | tail -1 | python -c "$show"
  echo -n "other:   "; AVLD_LIB_PATH=$other python bench.py $small 2>&1 | tail -1 | python -c "$show"
done
