"""DRAM traffic per pipeline stage (bytes per query) from a profiles/*_ncu_full_*_summary.csv.

  python tools/ncu_traffic.py profiles/r1d_ncu_full_c3_summary.csv profiles/r1d_traffic_c3.json
"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
def stage(name):
    if "k_traverse" in name: return "traverse"
    if "k_filter<0" in name or "k_clip<0" in name: return "clip_tet_tri"
    if "k_filter<1" in name or "k_clip<1" in name or "k_clip_slow" in name: return "clip_tet_tet"
    if "k_pack" in name: return "pack"
    if "k_surface" in name: return "surface"
    return None
out = {}
for r in rows[2:]:
    st = stage(r[0])
    if st is None:
        continue
    o = out.setdefault(st, {"dram_bytes": 0.0, "ncu_ms": 0.0, "launches": 0})
    o["dram_bytes"] += (float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]])
    o["ncu_ms"] += float(r[it]) * tscale[units[it]]
    o["launches"] += 1
json.dump({"source": sys.argv[1], "note": "dram__bytes_read.sum + dram__bytes_write.sum, summed over the launches of one query "
           "(ncu --set full --clock-control none; cold-cache, serialised)", "stages": out}, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
