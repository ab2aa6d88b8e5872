#!/bin/bash
for K in GIN GCN; do
KIND=$K timeout 600 python scripts/gin_probe.py 2>&1 | tail -1
KIND=$K timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gin_launches.csv python scripts/gin_probe.py > gpurun_out/gin_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/gin_launches.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
items=[(r[ik][:100], float(r[iv].replace(',',''))) for r in rows[1:]]
n=len(items)
last=items[-(n//7):]   # ~ one of the 6 iterations (+ setup)
tot=sum(v for _,v in last)
print(f"one fwd+bwd: {tot/1e6:.2f} ms over {len(last)} launches")
for k,v in last: 
    if v>100000: print(f"{v/1e6:7.3f} ms  {k}")
PY
done
