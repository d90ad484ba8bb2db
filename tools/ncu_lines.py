"""Top source lines of one kernel launch by warp-stall samples and executed instructions, over ALL source files
(an ncu report taken with --import-source on; kernels compiled with -lineinfo).

  python tools/ncu_lines.py report.ncu-rep 'regex:k_small' [function-substring] [top-n]
"""
import collections
import csv
import os
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
want_fn = sys.argv[3] if len(sys.argv) > 3 else ""
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 50
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "cuda,sass"],
                      capture_output=True, text=True).stdout
rows = list(csv.reader(text.splitlines()))
agg = collections.OrderedDict()
fn_seen, cur_file, cur_fn, hdr, use = [], None, None, None, False
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1])
    elif r[0] == "Function Name":
        cur_fn = r[1]
        if cur_fn not in fn_seen:
            fn_seen.append(cur_fn)
        use = (want_fn in cur_fn) and cur_fn == [f for f in fn_seen if want_fn in f][0]
    elif r[0] == "Line No":
        hdr = r
    elif use and hdr and len(r) == len(hdr) and r[0].strip().isdigit():
        iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        key = (cur_file, int(r[0]), r[1].strip()[:100])
        a = agg.setdefault(key, [0, 0, 0, collections.Counter()])
        try:
            a[0] += int(r[iS] or 0)
            a[1] += int(r[iI] or 0)
            a[2] += int(r[iT] or 0)
        except ValueError:
            continue
        for i, n in enumerate(hdr):
            if n.startswith("stall_") and "Not Issued" not in n:
                try:
                    a[3][n[6:]] += int(r[i] or 0)
                except ValueError:
                    pass
print("functions:", fn_seen, file=sys.stderr)
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
tott = sum(a[2] for a in agg.values()) or 1
print(f"samples {tot}, warp instructions {toti}, avg active lanes {tott / toti:.1f}")
byfile = collections.Counter()
for (f, ln, src), a in agg.items():
    byfile[f] += a[1]
print("instructions by file:", {k: f"{100 * v / toti:.1f}%" for k, v in byfile.items()})
for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    top = ", ".join(f"{k}={v}" for k, v in a[3].most_common(2))
    lanes = a[2] / a[1] if a[1] else 0
    print(f"{100 * a[1] / toti:5.1f}% inst {100 * a[0] / tot:5.1f}% smp {lanes:4.1f} ln  {f}:{ln:<5} {src}   [{top}]")
