#!/bin/bash
# tools/gpu_ncu_capture.sh <tag> "<config:scenes ...>": ncu captures of the final build: C3 (whole query), c3h (half-space kernels), C4 on 256 scenes, C5 on 16 scenes.
# The .ncu-rep files stay on the box (/tmp): only the summaries come back (gpurun_out is capped at 64 MiB).
tag=${1:-r2p}; specs=${2:-c3:16384 c3h:16384 c4:256 c5:16}
mkdir -p gpurun_out
K='k_small|k_clip|k_pack|k_surface|k_pairs|k_scan|k_traverse|k_filter|k_plane|k_leaf'
for spec in $specs; do
  cfg=${spec%%:*}; n=${spec##*:}
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches_$cfg.csv python tools/profile_one.py $spec 1 > gpurun_out/${tag}_launches_$cfg.log 2>&1
  timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$K" \
    -o /tmp/${tag}_full_$cfg -f python tools/profile_one.py $spec 1 > gpurun_out/${tag}_full_$cfg.log 2>&1
  python tools/ncu_summary.py /tmp/${tag}_full_$cfg.ncu-rep gpurun_out/${tag}_ncu_full_${cfg}_summary.csv
  python tools/ncu_kernels.py gpurun_out/${tag}_ncu_full_${cfg}_summary.csv gpurun_out/${tag}_full_$cfg.log $cfg gpurun_out/${tag}_kernels_$cfg.json $( [ "$cfg" = c4 ] || [ "$cfg" = c5 ] && echo $n ) | tail -2
  python tools/ncu_summary.py --launches gpurun_out/${tag}_launches_$cfg.csv gpurun_out/${tag}_launchlist_$cfg.csv | tail -1
done
# Hot source lines of the three dominant C3 kernels and the half-space kernels.
for k in "k_small:bool)0" "k_small:bool)1" "k_clip:bool)0" "k_clip:bool)1" "k_pack:"; do
  python tools/ncu_lines.py /tmp/${tag}_full_c3.ncu-rep "regex:${k%%:*}" "${k##*:}" 40 > "gpurun_out/${tag}_lines_c3_$(echo $k | tr -d ':)' | tr ' ' '_').txt" 2>&1
done
python tools/ncu_lines.py /tmp/${tag}_full_c3h.ncu-rep "regex:k_plane" "bool)0" 30 > gpurun_out/${tag}_lines_c3h_k_plane_count.txt 2>&1
python tools/ncu_lines.py /tmp/${tag}_full_c3h.ncu-rep "regex:k_plane" "bool)1" 30 > gpurun_out/${tag}_lines_c3h_k_plane_pack.txt 2>&1
du -sh gpurun_out; ls gpurun_out | head -40
