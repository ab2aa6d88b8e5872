#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/seg_probe3.py > gpurun_out/seg3.log 2>&1; echo "probe exit $?"; cat gpurun_out/seg3.log
SEG_PROBE_ITERS=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:seg_rows --csv --log-file gpurun_out/seg3_ncu.csv python scripts/seg_probe3.py > gpurun_out/seg3_ncu.log 2>&1; echo "ncu exit $?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/seg3_ncu.csv')) if len(r)>10]
hdr=rows[0]; 
i_id=hdr.index('ID'); i_k=hdr.index('Kernel Name'); i_m=hdr.index('Metric Name'); i_v=hdr.index('Metric Value')
d={}
for r in rows[1:]:
    d.setdefault(r[i_id],{})[r[i_m]]=r[i_v]; d[r[i_id]]['k']=r[i_k][:60]
for k,v in d.items(): print(k, v)
PY
