"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (name, grid) count / mean / total."""
import collections, csv, sys
lines = open(sys.argv[1]).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
agg = collections.OrderedDict()
for r in csv.DictReader(lines[start:]):
    if r["Metric Name"] == "gpu__time_duration.sum":
        agg.setdefault(r["Kernel Name"][:64] + " " + r["Grid Size"], []).append(float(r["Metric Value"]) / 1e3)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k:95s} n={len(v):4d} mean={sum(v)/len(v):8.1f} us  share={100*sum(v)/tot:5.1f}%")
print(f"total {tot/1e3:.2f} ms over {sum(len(v) for v in agg.values())} launches")
