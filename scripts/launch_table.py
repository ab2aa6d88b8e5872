"""Per-kernel table (ms, DRAM GB read / written, name) from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv` launch list: the kernels of the LAST iteration that ran (cold-cache, serialised times)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
kn, mn, mv, idc = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) > mv:
        d.setdefault(r[idc], {'name': r[kn]})[r[mn]] = float(r[mv].replace(',', ''))
L = list(d.values())
last = int(sys.argv[2]) if len(sys.argv) > 2 else 60
tot = 0.0
agg = collections.OrderedDict()
for k in L[-last:]:
    t = k.get('gpu__time_duration.sum', 0) / 1e6
    tot += t
    name = k['name'][:96]
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] += k.get('dram__bytes_read.sum', 0) / 1e9; a[3] += k.get('dram__bytes_write.sum', 0) / 1e9
print(f"# last {last} launches, {tot:.2f} ms in total; count, ms, DRAM GB read, GB written, kernel")
for name, a in agg.items():
    if a[1] >= 0.05:
        print(f"{a[0]:4d} {a[1]:8.3f} {a[2]:7.2f} {a[3]:7.2f}  {name}")
