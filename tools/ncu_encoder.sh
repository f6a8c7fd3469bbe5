#!/usr/bin/env bash
# ncu launch list (duration, tensor-pipe activity, grid) of one pass of the residual stand-in encoder (tools/encoder_probe.py)
cd "$(dirname "$0")/.."
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__grid_size --clock-control none \
  -k regex:"conv|gemm3|ew_|pool|gap|segment|latent" -s 100 -c 20 --csv --log-file gpurun_out/resenc.csv python tools/encoder_probe.py 2048 > gpurun_out/resenc.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/resenc.csv")) if len(r)>10]
h=rows[0]; ik,im,iv,ii=h.index("Kernel Name"),h.index("Metric Name"),h.index("Metric Value"),h.index("ID")
agg={}
for r in rows[1:]:
    agg.setdefault((int(r[ii]),r[ik][:46]),{})[r[im].split("__")[1][:14]]=r[iv]
tot=0
for k,m in sorted(agg.items()):
    print(k[0],k[1],m); tot+=float(m["time_duration."])
print("total us", tot/1e3)
PY
