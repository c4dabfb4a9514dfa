"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
iN, iV, iM = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iU = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or r[iM] != "gpu__time_duration.sum": continue
    v = float(r[iV].replace(",", ""))
    v = v / 1e3 if r[iU] in ("ns", "nsecond") else (v * 1e3 if r[iU] in ("ms", "msecond") else v)
    name = r[iN].split("(")[0][-60:]
    t, n = agg.get(name, (0.0, 0)); agg[name] = (t + v, n + 1)
tot = sum(t for t, _ in agg.values())
print(f"total {tot/1e3:.3f} ms over {sum(n for _, n in agg.values())} launches")
for name, (t, n) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}%  x{n:4d}  {t/n:9.1f} us/launch  {name}")
