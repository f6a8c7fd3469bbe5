cd /root/repo
small="--chunks 4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'convh|conv1|gemm3|ew_|pool' -c 40 --csv --log-file gpurun_out/convh_launches.csv python bench.py $small > gpurun_out/ncu_convh.log 2>&1; echo rc $?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/convh_launches.csv')) if len(r)>10]
hdr=rows[0]; i_k=hdr.index('Kernel Name'); i_m=hdr.index('Metric Name'); i_v=hdr.index('Metric Value'); i_id=hdr.index('ID')
agg={}
for r in rows[1:]:
    agg.setdefault((int(r[i_id]),r[i_k][:60]),{})[r[i_m]]=r[i_v]
for (i,k),m in sorted(agg.items())[:24]:
    print(i,k,m)
PY
