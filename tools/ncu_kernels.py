"""Per-kernel-group DRAM bytes, warp instructions and occupancy figures of ONE query, from a profiles/*_summary.csv
(tools/ncu_summary.py over an `ncu --set full --clock-control none` capture of tools/profile_one.py <config>:<scenes> 1),
in the form bench.py reads for `roofline.traffic`, `roofline.issue_slots` and `roofline.compulsory_dram`.

  python tools/ncu_kernels.py profiles/r2_ncu_full_c3_summary.csv gpurun_out/r2x_full.log c3 profiles/r2_kernels_c3.json [scenes]

The log is profile_one.py's stdout: its last line is the stats dict of the profiled query, from which the compulsory
DRAM traffic is computed (every byte that must cross the HBM interface at least once however good the caches are):
poses in + survivor lists written and read + staged polygons written and read + the packed result written; the mesh
assets (KBs to MBs, instanced by every scene) are counted once.
"""
import ast
import csv
import json
import sys

summary, log, config, out_path = sys.argv[1:5]
scenes = int(sys.argv[5]) if len(sys.argv) > 5 else None   # scenes of the captured query (None: the config's default size)
rows = list(csv.reader(open(summary)))
hdr, units = rows[0], rows[1]
col = {n: hdr.index(n) for n in hdr}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def group(name):
    if "k_small_pairs<0" in name or "k_filter<0" in name: return "mid_phase_tet_tri"
    if "k_small_pairs<1" in name or "k_filter<1" in name: return "mid_phase_tet_tet"
    if "k_leaf_undecided" in name: return "mid_phase_tet_tet"   # both kinds; timed with the tet-tet stage (bench.py marks)
    if "k_traverse" in name: return "traverse"
    if "k_clip<0" in name: return "exact_clip_tet_tri"
    if "k_clip<1" in name or "k_clip_slow" in name: return "exact_clip_tet_tet"
    if "k_pack" in name or "k_plane<1" in name: return "pack"
    if "k_plane" in name: return "half_space"
    if "k_surface" in name: return "surface"
    return "other"


def val(r, name):
    i = col[name]
    u = units[i]
    return float(r[i]) * scale.get(u, 1.0)


groups, total_dram, total_ms = {}, 0.0, 0.0
for r in rows[2:]:
    g = groups.setdefault(group(r[0]), {"kernels": [], "dram_bytes": 0.0, "ncu_ms": 0.0, "warp_instructions": 0.0, "launches": 0,
                                        "_w": 0.0, "issue_active_pct": 0.0, "warps_active_pct": 0.0, "active_lanes": 0.0,
                                        "fp64_pipe_pct": 0.0})
    ms = float(r[col["gpu__time_duration.sum"]]) * tscale[units[col["gpu__time_duration.sum"]]]
    dram = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    g["dram_bytes"] += dram
    g["ncu_ms"] += ms
    g["launches"] += 1
    g["warp_instructions"] += float(r[col["smsp__inst_executed.sum"]])
    if r[0] not in g["kernels"]:
        g["kernels"].append(r[0])
    # time-weighted averages of the per-launch percentages
    g["_w"] += ms
    g["issue_active_pct"] += ms * float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]])
    g["warps_active_pct"] += ms * float(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]])
    g["active_lanes"] += ms * float(r[col["smsp__thread_inst_executed_per_inst_executed.ratio"]])
    g["fp64_pipe_pct"] += ms * float(r[col["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]])
    total_dram += dram
    total_ms += ms
for g in groups.values():
    w = g.pop("_w") or 1.0
    for k in ("issue_active_pct", "warps_active_pct", "active_lanes", "fp64_pipe_pct"):
        g[k] = g[k] / w
stats = None
for line in reversed(open(log).read().splitlines()):
    if line.startswith("{") and "tet_tri_pairs" in line:
        stats = ast.literal_eval(line)
        break
compulsory = staging = None
if stats:
    # Compulsory: what any implementation must move -- poses in, the packed result out (the instanced mesh assets are KBs
    # to MBs, read once). Staging: this design's own intermediate lists, each written once and read once -- survivor
    # entries, staged polygons, and (per-level pipeline only) the leaf lists.
    surv = stats["tet_tri_full_clips"] + stats["tet_tet_full_clips"]
    soup = stats["polygons"] * 40 + stats["vertices"] * 32
    result = stats["vertices"] * 32 + stats["polygons"] * 108 + (stats["polygons"] + stats["vertices"]) * 4
    poses = stats["geometry_pairs_tested"] * 2 * 96
    compulsory = poses + result
    generic = "traverse" in groups
    staging = 2 * 16 * surv + 2 * soup + (2 * 16 * stats["leaf_pairs"] if generic else 0)
out = {"config": config, "scenes": scenes, "source": summary, "sms": 148,
       "note": "one query under `ncu --set full --clock-control none` (cold caches, kernels serialised): shares and counters, "
               "not timings",
       "groups": groups, "dram_bytes_whole_query": total_dram, "ncu_ms_whole_query": total_ms,
       "compulsory_dram_bytes": compulsory, "staging_dram_bytes": staging, "stats": stats}
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps({k: {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk != "kernels"}
                  for k, v in groups.items()}, indent=1))
print("whole query: dram", total_dram / 1e9, "GB; compulsory", (compulsory or 0) / 1e9, "GB; staging", (staging or 0) / 1e9,
      "GB; ncu ms", total_ms)
