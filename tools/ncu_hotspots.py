"""Markdown hot-spot table from an `ncu --page source --csv --print-source cuda,sass` export (plain or .gz): per source-file share of the
stall samples, the top source lines, and the executed-instruction mix by SASS opcode (each SASS row counted once, under the file
section it is listed in first).
    python tools/ncu_hotspots.py profiles/r02d_k_step_source.csv.gz [top] > profiles/r02d_k_step_hotspots.md"""
import collections, csv, gzip, io, re, sys

path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
cur = None; hdr = None; func = None
lines = collections.OrderedDict(); ops = collections.Counter(); seen = set()
for row in csv.reader(io.StringIO(raw)):
    if not row: continue
    if row[0] == "File Path": cur = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": func = row[1]; continue
    if row[0] == "Line No": hdr = row; iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); continue
    if hdr is None: continue
    if row[0] != "" and row[2] == "-":
        a = lines.setdefault((cur, int(row[0])), [0, 0, row[1].strip(), collections.Counter()])
        a[0] += int(row[iS]); a[1] += int(row[iI])
        for k in hdr:
            if k.startswith("stall_") and "Not Issued" not in k:
                try: a[3][k[6:]] += int(row[hdr.index(k)])
                except ValueError: pass
    elif row[0] == "" and len(row) > iI and row[2].startswith("0x") and row[2] not in seen:
        seen.add(row[2])
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", row[3])
        if m: ops[m.group(2).split(".")[0]] += int(row[iI])
ts = sum(a[0] for a in lines.values()); ti = sum(ops.values())
print(f"# Hot spots of `{func}`\n\nSource: `{path}` ({ts} stall samples, {ti / 1e6:.1f} M warp-instructions executed per launch).\n")
byfile = collections.Counter()
for (f, l), a in lines.items(): byfile[f] += a[0]
print("| file | share of stall samples |\n|---|---|")
for f, v in byfile.most_common():
    if v / ts > 0.002: print(f"| `{f}` | {100 * v / ts:.1f} % |")
print(f"\n| samples | line | top stall reasons | source |\n|---|---|---|---|")
for (f, l), a in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{k} {v}" for k, v in a[3].most_common(2))
    print(f"| {100 * a[0] / ts:.2f} % | `{f}:{l}` | {st} | `{a[2][:90].replace('|', '/')}` |")
print(f"\n| SASS opcode | share of executed warp-instructions |\n|---|---|")
for k, v in ops.most_common(16): print(f"| {k} | {100 * v / ti:.1f} % |")
