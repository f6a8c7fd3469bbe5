#!/usr/bin/env bash
# what bounds convh_kernel: per-layer launch times (ncu launch list) with the timing probes of AVLD_CONVH_DBG
cd "$(dirname "$0")/.."
export AVLD_LIB_PATH=$PWD/amphibian_vae_latent_detector_b200/libavld_bringup.so   # the probes live in the bring-up build only
for d in 0 2; do
  AVLD_CONVH_DBG=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:convh -c 6 --csv --log-file gpurun_out/convh_dbg$d.csv \
    python bench.py --chunks 2048 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
  echo "dbg $d: $(python tools/ncu_summary.py launches gpurun_out/convh_dbg$d.csv | grep convh | awk '{print $(NF-1)}' | tr '\n' ' ')"
done
