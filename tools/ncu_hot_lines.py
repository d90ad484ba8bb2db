"""Top source lines of one kernel by warp-stall samples, from an ncu report taken with --import-source on.

  python tools/ncu_hot_lines.py report.ncu-rep 'regex:k_clip' [launch-index] [top-n]
"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "cuda,sass"],
                      capture_output=True, text=True).stdout
rows = list(csv.reader(text.splitlines()))
# Split into per-launch tables: a table starts at a header row whose first cell is "Line No".
tables, cur = [], None
for r in rows:
    if r and r[0] == "Line No":
        cur = {"hdr": r, "rows": []}
        tables.append(cur)
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
tables = [t for t in tables if len(t["rows"]) > 50]
print(f"{len(tables)} tables; using #{which}", file=sys.stderr)
t = tables[which]
h = t["hdr"]
iS, iI = h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.OrderedDict()
for r in t["rows"]:
    key = (r[0], r[1].strip()[:110])
    a = agg.setdefault(key, [0, 0, collections.Counter()])
    try:
        a[0] += int(r[iS] or 0); a[1] += int(r[iI] or 0)
    except ValueError:
        continue
    for i in stall_cols:
        try:
            a[2][h[i]] += int(r[i] or 0)
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print(f"total samples {tot}, warp instructions {toti}")
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    top = ", ".join(f"{k[6:]}={v}" for k, v in a[2].most_common(3))
    print(f"{100*a[0]/tot:5.1f}% smp {100*a[1]/toti:5.1f}% inst  L{ln:>5} {src}   [{top}]")
