#!/bin/bash
# ncu captures of the final build: C3 (whole query), c3h (half-space kernels), C4 on 256 scenes, C5 on 16 scenes.
mkdir -p gpurun_out
K='k_small|k_clip|k_pack|k_surface|k_pairs|k_scan|k_traverse|k_filter|k_plane'
for spec in c3:16384 c3h:16384 c4:256 c5:16; do
  cfg=${spec%%:*}
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2p_launches_$cfg.csv python tools/profile_one.py $spec 1 > gpurun_out/r2p_launches_$cfg.log 2>&1
  timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$K" \
    -o gpurun_out/r2p_full_$cfg -f python tools/profile_one.py $spec 1 > gpurun_out/r2p_full_$cfg.log 2>&1
  tail -1 gpurun_out/r2p_full_$cfg.log | cut -c1-200
done
ls -la gpurun_out/r2p_*.ncu-rep
