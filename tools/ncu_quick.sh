#!/usr/bin/env bash
# per-launch duration / DRAM bytes / L2 hit rate of the kernels matching a regex (serialised, cold: shares, not absolutes)
# Usage: tools/ncu_quick.sh '<kernel regex>' [count] [VAR=1 ... for the bring-up build]
cd "$(dirname "$0")/.."
re=$1; cnt=${2:-12}; shift 2
envs=""; [ $# -gt 0 ] && envs="env AVLD_LIB_PATH=amphibian_vae_latent_detector_b200/libavld_bringup.so $*"
timeout 600 $envs ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"$re" -s 2 -c $cnt --csv --log-file gpurun_out/quick.csv python bench.py --chunks 4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/quick.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/quick.csv')) if len(r)>10]
h=rows[0]; ik,im,iv,ii=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value'),h.index('ID')
agg={}
for r in rows[1:]:
    agg.setdefault((int(r[ii]),r[ik][:48]),{})[r[im].replace('.sum','').replace('.avg.pct_of_peak_sustained_active','').split('__',1)[1][:26]]=r[iv]
for k,m in sorted(agg.items()): print(k[0],k[1],m)
PY
