"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python scripts/show_launches.py file.csv [steps]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    e = agg.setdefault(d["Kernel Name"][:70], [0, 0.0])
    e[0] += 1
    e[1] += v
for k, (c, t) in agg.items():
    print(f"{k:72s} {c:4d} {t / steps:10.1f} us/step")
print("total/step", sum(v[1] for v in agg.values()) / steps)
