"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line (samples, warp instructions)."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None
agg = collections.OrderedDict()
for row in csv.reader(open(path)):
    if not row: continue
    if row[0] == "File Path": cur = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; continue
    if row[0] == "" or hdr is None: continue
    if row[2] != "-": continue
    d = dict(zip(hdr, row))
    key = (cur, int(row[0]))
    s = int(row[hdr.index("# Samples")]); i = int(row[hdr.index("Instructions Executed")])
    a = agg.setdefault(key, [0, 0, row[1], collections.Counter()])
    a[0] += s; a[1] += i
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            try: a[3][k] += int(row[hdr.index(k)])
            except ValueError: pass
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
print(f"total samples {ts}  total warp-instructions {ti}")
byfile = collections.Counter()
for (f, l), a in agg.items(): byfile[f] += a[0]
print({k: f"{100*v/ts:.1f}%" for k, v in byfile.items()})
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ",".join(f"{k[6:]}={v}" for k, v in a[3].most_common(3))
    print(f"{100*a[0]/ts:5.2f}% smp {100*a[1]/ti:5.2f}% ins  {f}:{l:<4d} {st:50s} | {a[2].strip()[:110]}")
