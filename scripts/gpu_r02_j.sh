#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PATHS=fused timeout 200 python scripts/tconv_probe.py > gpurun_out/r02j_plain.log 2>&1 &&
PATHS=fused timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02j_tconv_launches.csv python scripts/tconv_probe.py > gpurun_out/r02j_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02j_tconv_launches.csv')) if len(r)>10]
hdr=rows[0]
i_id=hdr.index('ID'); i_k=hdr.index('Kernel Name'); i_m=hdr.index('Metric Name'); i_v=hdr.index('Metric Value')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[i_id],{'k':r[i_k][:80]})[r[i_m]]=float(r[i_v].replace(',',''))
items=list(d.values())
for it in items[-60:]:
    if it.get('gpu__time_duration.sum',0) > 200000:
        print(f"{it['gpu__time_duration.sum']/1e6:8.3f} ms  rd {it.get('dram__bytes_read.sum',0)/1e9:6.2f} GB wr {it.get('dram__bytes_write.sum',0)/1e9:6.2f} GB  {it['k']}")
PY
