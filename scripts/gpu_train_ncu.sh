#!/bin/bash
mkdir -p gpurun_out
STEPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python scripts/train_probe.py > gpurun_out/train_ncu.log 2>&1; echo "ncu exit $?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/train_launches.csv')) if len(r)>10]
hdr=rows[0]
i_id=hdr.index('ID'); i_k=hdr.index('Kernel Name'); i_v=hdr.index('Metric Value')
items=[(r[i_k], float(r[i_v].replace(',',''))) for r in rows[1:]]
n=len(items)//3          # 3 steps (2 warm + 1): take the last third
last=items[-n:]
agg=collections.Counter(); cnt=collections.Counter()
for k,v in last: agg[k[:80]]+=v; cnt[k[:80]]+=1
tot=sum(agg.values())
print(f"one step: {tot/1e6:.2f} ms over {len(last)} launches")
for k,v in agg.most_common(24): print(f"{v/1e6:8.3f} ms {v/tot:6.1%} x{cnt[k]:3d}  {k}")
PY
