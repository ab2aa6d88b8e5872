#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/gat_launches.csv python scripts/gat_probe.py > gpurun_out/gat_ncu.log 2>&1; echo "ncu exit $?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/gat_launches.csv')) if len(r)>10]
hdr=rows[0]
i_id=hdr.index('ID'); i_k=hdr.index('Kernel Name'); i_m=hdr.index('Metric Name'); i_v=hdr.index('Metric Value')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[i_id],{'k':r[i_k][:90]})[r[i_m]]=float(r[i_v].replace(',',''))
# print the launches of the last aggregate fwd+bwd and of one fwd
items=list(d.values())
for it in items[-140:]:
    if it.get('gpu__time_duration.sum',0) > 50000:
        print(f"{it['gpu__time_duration.sum']/1e6:8.3f} ms  rd {it.get('dram__bytes_read.sum',0)/1e9:6.2f} GB wr {it.get('dram__bytes_write.sum',0)/1e9:6.2f} GB  {it['k']}")
PY
